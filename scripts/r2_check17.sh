mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or packed_grids or path_records" 2>&1 | tail -2
for ch in 0 296 1036; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-dubins --no-collision --no-cpu --no-strong --no-informed --no-class-api --e2e-chunk $ch 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $ch value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'ms', round(d['e2e']['ms_per_step'],2), 'trees', round(d['e2e']['trees_mode']['value']), d['e2e']['matches_device_arm'])"
done
