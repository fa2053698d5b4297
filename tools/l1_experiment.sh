run() { python bench.py --steps 100 --warmup 10 --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['config']; print('$1', c['threads_per_instance'], c['instances_per_cta'], c['grid_ctas'], c['smem_bytes_per_cta'], 'value=%.0f kernel_ms=%.4f'%(d['value'], d['roofline']['kernel_ms']))"; }
export RESCO_B200_GROUP=4 RESCO_B200_BLOCK=64
RESCO_B200_SMEM_EXTRA=25000 run "G4 1cta/sm L1~124K"
RESCO_B200_SMEM_EXTRA=90000 run "G4 1cta/sm L1~60K"
RESCO_B200_SMEM_EXTRA=0 run "G4 2cta/sm L1~60K"
export RESCO_B200_GROUP=2 RESCO_B200_BLOCK=64
RESCO_B200_SMEM_EXTRA=10000 run "G2 1cta/sm L1~190K"
RESCO_B200_SMEM_EXTRA=135000 run "G2 1cta/sm L1~60K"
