#!/bin/bash
# build the instrumented library (per-phase cycle accounting) next to the product library
cd "$(dirname "$0")/.."
mkdir -p ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC -DRS_PHASE_CLOCKS=1 \
  -o ab/clocks.so resco_b200/csrc/sim.cu && echo "built ab/clocks.so; run: RESCO_B200_LIB=\$PWD/ab/clocks.so python tools/phase_clocks.py"
