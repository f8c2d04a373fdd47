timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
RRTK_PLAN_IMPL=grid timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_fullsize.py -q -m gpu -k "not takes_the_wide_kernel" 2>&1 | tail -2
RRTK_PLAN_IMPL=scan timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu -k "not takes_the_wide_kernel" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-informed --no-class-api --no-collision --no-dubins --no-cpu 2>gpurun_out/b33.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'trees', round(d['e2e']['trees_mode']['value']), 'overlapped', round(d['overlapped_steps']['value']), 'strong', round(d['strong_scaling']['value']), d['roofline']['frac'], d['roofline'].get('kernel'))
"
tail -2 gpurun_out/b33.err
