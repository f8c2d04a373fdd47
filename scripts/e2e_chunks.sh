# end-to-end plans/s for several pipeline chunk sizes of rrtk_ctx_plan_worlds
for c in "$@"; do
  python bench.py --steps 3 --warmup 3 --no-cpu --no-collision --no-dubins --e2e-chunk $c 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $c device', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['e2e']['ms_per_step'],1))"
done
