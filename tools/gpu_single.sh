#!/bin/bash
# single-buffer tile experiment on the big maps
cd "$(dirname "$0")/.."
run() { python bench.py "${@:2}" --steps 20 --warmup 3 --preroll 60 --no-cpu 2>&1 | tail -1 | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read()); c=d['config']; print('$1', c['threads_per_instance'], c['instances_per_cta'], c['grid_ctas'], c['smem_bytes_per_cta'], c['tile_buffers'], 'value=%.0f e2e=%.0f kernel_ms=%.4f'%(d['value'], d['e2e']['value'], d['roofline']['kernel_ms']))
except Exception as e: print('$1 FAILED', e)"; }
ING="--map ingolstadt21 --vcap 1024 --n-env 2048"
GRID="--map grid4x4 --synthetic-rate 600 --vcap 1024 --n-env 2048"
run ing21-auto $ING
RESCO_B200_SINGLE=0 run ing21-double $ING
RESCO_B200_SINGLE=1 RESCO_B200_GROUP=2 run ing21-single-G2 $ING
RESCO_B200_SINGLE=1 RESCO_B200_GROUP=1 RESCO_B200_REGCAP=0 run ing21-single-1cta $ING
run grid-auto $GRID
RESCO_B200_SINGLE=0 run grid-double $GRID
RESCO_B200_SINGLE=1 RESCO_B200_GROUP=1 run ing21-single-2cta $ING
RESCO_B200_SINGLE=1 RESCO_B200_GROUP=1 run grid-single-2cta $GRID
