# device-arm plans/s of experiment builds exp_<name>.so ("main" = the in-tree library):  scripts/variants.sh T name...
T=$1; shift
for v in "$@"; do
  L=$PWD/exp_$v.so; [ "$v" = main ] && L=$PWD/rrtplanner_b200/librrtk.so
  RRTK_LIB=$L timeout 300 python bench.py --steps 3 --warmup 3 --plans ${PLANS:-4096} --threads $T --plan-only 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v T=$T plans/s', round(d['value']), 'plan_ms', round(d['kernel_ms']['plan'],2), 'blocks/SM', d['roofline']['blocks_per_sm'])
except Exception as e: print('variant $v failed', e)"
done
