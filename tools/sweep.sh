#!/bin/bash
# usage: tools/sweep.sh  -- throughput sweep over launch parameters / library variants (run under gpurun)
cd "$(dirname "$0")/.."
for lib in resco_b200/csrc/libresco_b200.so $(ls resco_b200/csrc/variants/*.so 2>/dev/null); do
for b in ${BLOCKS:-64 128}; do for r in ${REGCAPS:-1}; do for vcap in ${VCAPS:-128}; do for pe in ${PERSIST:-1}; do
  out=$(RESCO_B200_LIB=$PWD/$lib RESCO_B200_BLOCK=$b RESCO_B200_REGCAP=$r RESCO_B200_PERSIST=$pe python bench.py --steps ${STEPS:-60} --warmup 5 --no-cpu --vcap $vcap 2>&1 | tail -1)
  echo "$(basename $lib) block=$b regcap=$r vcap=$vcap persist=$pe $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("value=%.0f e2e=%.0f kernel_ms=%.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"]))
except Exception as e: print("ERR", e)')"
done; done; done; done; done
