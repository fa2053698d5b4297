#!/bin/bash
# usage: CONFIGS="tpi:group:regcap ..." tools/sweep.sh  -- throughput sweep over launch shapes (run under gpurun)
cd "$(dirname "$0")/.."
for cfg in ${CONFIGS:-128:1:1}; do
  IFS=: read b g r <<< "$cfg"
  for vcap in ${VCAPS:-128}; do
  out=$(RESCO_B200_BLOCK=$b RESCO_B200_GROUP=$g RESCO_B200_REGCAP=$r python bench.py --steps ${STEPS:-60} --warmup 5 --no-cpu --vcap $vcap ${EXTRA} 2>&1 | tail -1)
  echo "tpi=$b group=$g regcap=$r vcap=$vcap $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("value=%.0f e2e=%.0f kernel_ms=%.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"]))
except Exception as e: print("ERR", e)')"
  done
done
