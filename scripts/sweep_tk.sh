# device-arm plans/s for block shapes given as "T:K" pairs (RRTK_PLAN_K picks the samples per round)
for cfg in "$@"; do
  T=${cfg%%:*}; K=${cfg##*:}
  RRTK_PLAN_K=$K timeout 300 python bench.py --steps 3 --warmup 3 --plans ${PLANS:-2960} --threads $T --plan-only 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('T=$T K=$K plans/s', round(d['value']), 'plan_ms', round(d['kernel_ms']['plan'],2), 'blocks/SM', d['roofline']['blocks_per_sm'], 'frac', round(d['roofline']['frac'],3))
except Exception as e: print('T=$T K=$K failed', e)"
done
