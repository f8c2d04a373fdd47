python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -3
python bench.py --informed-only --steps 3 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('informed', round(d['plans_per_s']), d['ms_per_launch'], d.get('blocks_per_sm'), 'e2e', round(d['e2e']['value']), d.get('matches_oracle'), d['roofline']['frac'])
"
