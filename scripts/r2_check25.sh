python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-strong --no-informed --no-class-api --no-collision --no-dubins --no-cpu 2>gpurun_out/b25.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'trees', round(d['e2e']['trees_mode']['value']), d['e2e'].get('same_results_as_device_arm'), d['e2e']['trees_mode'].get('same_results_as_device_arm'))
"
tail -2 gpurun_out/b25.err
