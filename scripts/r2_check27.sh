for L in rrtplanner_b200/librrtk.so exp_mb2.so; do
RRTK_LIB=$PWD/$L python bench.py --informed-only --no-cpu --steps 3 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$L', round(d['plans_per_s']), d['ms_per_launch'], d.get('blocks_per_sm'), 'e2e', round(d['e2e']['value']))
"
done
