#!/bin/bash
# end-of-round evidence run on one B200 (one gpurun call): smoke, GPU tests, the bench lines of all four configs, the CPU
# reference arm, ncu launch lists + one full capture of the env-step kernel per config
cd "$(dirname "$0")/.."
T=${TAG:-r02v}
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${T}_pytest_gpu.log; cat $O/${T}_pytest_gpu.log
python bench.py > $O/${T}_bench_c2.json 2> $O/${T}_bench.err; cut -c1-160 $O/${T}_bench_c2.json
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_c2_driver_window.json 2>> $O/${T}_bench.err; cut -c1-160 $O/${T}_bench_c2_driver_window.json
python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference_arm.json 2>> $O/${T}_bench.err; cut -c1-200 $O/${T}_bench_reference_arm.json
for c in c3 c4 c5; do
  python bench.py --config $c --steps 40 --warmup 5 > $O/${T}_bench_$c.json 2>> $O/${T}_bench.err; cut -c1-160 $O/${T}_bench_$c.json
done
TAG=$T tools/gpu_profile_r02.sh c2 c3 c5
