#!/bin/bash
# launch-shape sweep on the GPU box for the big-tile maps (one instance per CTA): threads per instance 64..512
cd "$(dirname "$0")/.."
for cfg in "ingolstadt21 0" "grid4x4 600"; do
set -- $cfg
for b in ${BLOCKS:-64 256 512}; do
RESCO_B200_BLOCK=$b python bench.py --map $1 --synthetic-rate $2 --vcap ${VCAP:-1024} --n-env ${NENV:-2048} --steps ${STEPS:-20} --warmup 3 --preroll ${PREROLL:-60} --no-cpu 2>&1 | tail -1 | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); c=d["config"]; print(sys.argv[1], sys.argv[2], "tpi", c["threads_per_instance"], "G", c["instances_per_cta"], "value=%.0f e2e=%.0f kernel_ms=%.3f vbar=%.1f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["vbar_active_vehicles"]))
except Exception as e: print(sys.argv[1], sys.argv[2], "FAILED", e)' $1 $b
done
done
