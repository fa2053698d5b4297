#!/bin/bash
# quick A/B on the GPU box: parity on two maps (tools/gpu_quick.py) with the in-tree library, then the C2 bench line
# (kernel-only numbers) for the in-tree library and for every ab/*.so (RESCO_B200_LIB), interleaved
cd "$(dirname "$0")/.."
python tools/gpu_quick.py 2>&1 | tail -4
for i in 1 2; do
for lib in resco_b200/csrc/libresco_b200.so ab/*.so; do
[ -f "$lib" ] || continue
RESCO_B200_LIB=$PWD/$lib python bench.py --steps ${STEPS:-300} --warmup 10 --no-cpu ${EXTRA} 2>/dev/null | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.0f e2e=%.0f kernel_ms=%.4f vbar=%.1f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["vbar_active_vehicles"]))' $lib
done
done
