# GPU-box check: parity tests, then a short device-arm bench per configuration given as "T:K" pairs
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for cfg in "$@"; do
  T=${cfg%%:*}; K=${cfg##*:}
  RRTK_PLAN_K=$K timeout 300 python bench.py --steps 2 --warmup 3 --plans ${PLANS:-2368} --threads $T --plan-only 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('T=$T K=$K', round(d['value']), d['kernel_ms']['plan'], d['roofline']['blocks_per_sm'], round(d['roofline']['frac'],3))
except Exception as e: print('T=$T K=$K failed', e)"
done
