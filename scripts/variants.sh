# device-arm plans/s of experiment builds exp_<name>.so:  scripts/variants.sh T name...
T=$1; shift
for v in "$@"; do
  RRTK_LIB=$PWD/exp_$v.so timeout 300 python bench.py --steps 3 --warmup 3 --plans ${PLANS:-2960} --threads $T --no-e2e --no-cpu --no-collision 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v T=$T plans/s', round(d['value']), 'plan_ms', round(d['kernel_ms']['plan'],2))
except Exception as e: print('variant $v failed', e)"
done
