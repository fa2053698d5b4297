#!/bin/bash
# usage: tools/sweep.sh  -- throughput sweep over launch parameters (run under gpurun)
cd "$(dirname "$0")/.."
for b in ${BLOCKS:-64 128}; do for r in ${REGCAPS:-0 1}; do for vcap in ${VCAPS:-256}; do for pe in ${PERSIST:-0 1}; do
  out=$(RESCO_B200_BLOCK=$b RESCO_B200_REGCAP=$r RESCO_B200_PERSIST=$pe python bench.py --steps ${STEPS:-40} --warmup 5 --no-cpu --vcap $vcap 2>&1 | tail -1)
  echo "block=$b regcap=$r vcap=$vcap persist=$pe $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("value=%.0f e2e=%.0f kernel_ms=%.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"]))
except Exception as e: print("ERR", e)')"
done; done; done; done
