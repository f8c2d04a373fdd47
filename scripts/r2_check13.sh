mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/variants.sh 0 prev main 2>&1 | tail -2
RRTK_LIB=$PWD/exp_clk.so timeout 300 python scripts/phase_clocks.py 1036 128 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-dubins --no-informed --no-class-api --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'trees', round(d['e2e']['trees_mode']['value']), 'strong', round(d['strong_scaling']['value']), 'cc', d['collision_microbench']['ms_per_launch'], d['collision_microbench']['roofline']['frac'])"
