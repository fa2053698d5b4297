#!/bin/bash
# A/B of the in-tree library against every ab/*.so on the C2 bench shape (kernel-only numbers; no parity run:
# the ab/ builds may be deliberately non-equivalent cost-attribution experiments).  Build variants here first, e.g.
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC \
#        -DRS_BALANCED_PLAN=1 -o ab/balanced.so resco_b200/csrc/sim.cu
# and pass EXTRA='--map ingolstadt21 --vcap 1024 --n-env 2048' for the big-map launch shape.
cd "$(dirname "$0")/.."
for i in 1 2; do
for lib in resco_b200/csrc/libresco_b200.so ab/*.so; do
[ -f "$lib" ] || continue
RESCO_B200_LIB=$PWD/$lib python bench.py --steps ${STEPS:-200} --warmup 5 --no-cpu ${EXTRA} 2>/dev/null | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.0f e2e=%.0f kernel_ms=%.4f vbar=%.1f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["vbar_active_vehicles"]))' $lib
done
done
