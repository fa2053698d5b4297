#!/bin/bash
# multi-GPU bench lines on one box (gpurun --gpus N): the driver's launch line for every config that shards
cd "$(dirname "$0")/.."
N=${N:-8}; T=${TAG:-r02w}; O=gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 20 --warmup 5 > $O/${T}_bench_c2_${N}gpu.json 2> $O/${T}_scale.err; cut -c1-200 $O/${T}_bench_c2_${N}gpu.json
run --config c4 --steps 40 --warmup 5 > $O/${T}_bench_c4_${N}gpu.json 2>> $O/${T}_scale.err; cut -c1-200 $O/${T}_bench_c4_${N}gpu.json
run --config c5 --steps 40 --warmup 5 > $O/${T}_bench_c5_${N}gpu.json 2>> $O/${T}_scale.err; cut -c1-200 $O/${T}_bench_c5_${N}gpu.json
run --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference_${N}gpu.json 2>> $O/${T}_scale.err; cut -c1-200 $O/${T}_bench_reference_${N}gpu.json
tail -3 $O/${T}_scale.err
