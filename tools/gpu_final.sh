#!/bin/bash
# end-of-round evidence run (one gpurun call): smoke, GPU tests, the C2 bench line, the CPU reference arm, the ncu
# launch list of the bench command, one full ncu capture of k_run, and the C3-like / C5-like big-map lines
cd "$(dirname "$0")/.."
T=${TAG:-r01j}
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/${T}_pytest_gpu.log; cat $O/${T}_pytest_gpu.log
python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench.err; cut -c1-200 $O/${T}_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference_arm.json 2>> $O/${T}_bench.err; cut -c1-300 $O/${T}_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_ncu.csv \
    python bench.py --steps 12 --warmup 3 --preroll 60 --no-cpu > $O/${T}_ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_run -s 50 -c 1 -f -o $O/${T}_krun \
    python bench.py --steps 12 --warmup 3 --preroll 60 --no-cpu > $O/${T}_ncu.log 2>&1
python bench.py --map ingolstadt21 --vcap 1024 --n-env 8192 --steps 20 --warmup 3 --preroll 60 --no-cpu > $O/${T}_bench_c3.json 2>> $O/${T}_bench.err; cut -c1-160 $O/${T}_bench_c3.json
python bench.py --map grid4x4 --synthetic-rate 600 --vcap 1024 --n-env 16384 --steps 20 --warmup 3 --preroll 60 --no-cpu > $O/${T}_bench_c5.json 2>> $O/${T}_bench.err; cut -c1-160 $O/${T}_bench_c5.json
ls -la $O | tail -12
