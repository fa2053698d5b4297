#!/bin/bash
# instances-per-CTA sweep on the C2 shape (wave quantisation of the persistent grid: 4096 instances = 512 groups of 8
# on 148 CTAs is 3.46 rounds; 586 groups of 7 is 3.96)
cd "$(dirname "$0")/.."
run() { python bench.py "${@:2}" --steps ${STEPS:-200} --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read()); c=d['config']; print('$1', c['threads_per_instance'], c['instances_per_cta'], c['grid_ctas'], c['smem_bytes_per_cta'], 'value=%.0f e2e=%.0f sync=%.0f kernel_ms=%.4f'%(d['value'], d['e2e']['value'], d['e2e']['sync_value'], d['roofline']['kernel_ms']))
except Exception as e: print('$1 FAILED', e)"; }
for g in 8 7 6 5; do RESCO_B200_GROUP=$g run "n4096-G$g"; done
for g in 8 7; do RESCO_B200_GROUP=$g run "n2048-G$g" --n-env 2048; done
for g in 8 7; do RESCO_B200_GROUP=$g run "n4144-G$g" --n-env 4144; done
