mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed_grids or path_records" 2>&1 | tail -5
timeout 600 python scripts/lpt_test.py 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --no-dubins > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -3 gpurun_out/r2_bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench1.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e trees', round(d['e2e']['trees_mode']['value']), 'match', d['e2e']['matches_device_arm'], d['e2e']['trees_mode']['matches_device_arm'], d['e2e']['host_packer_matches_device_packer'])
print('roofline frac', round(d['roofline']['frac'],3), 'peak', round(d['roofline']['peak']), 'matches_oracle', d.get('matches_oracle'))
print('strong', d.get('strong_scaling'))
c=d['collision_microbench']; print('cc', c['ms_per_launch'], c['roofline']['frac'], c['bit_grid_kernel']['ms_per_launch'], c['bit_grid_kernel']['roofline']['frac'], c.get('matches_oracle'))
PY
