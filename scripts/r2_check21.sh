python -m pytest tests/test_gpu_clearance.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -8
python bench.py --collision-only 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
for k in ('kernel', 'ms_per_launch', 'field_build_ms', 'same_outputs_as_bit_grid_kernel', 'matches_oracle'): print(k, d.get(k))
print('frac', d['roofline']['frac'], 'iso', d['isotropic_field_kernel']['ms_per_launch'], d['isotropic_field_kernel']['roofline']['frac'], 'build', d['isotropic_field_kernel']['field_build_ms'])
"
for p in 32 64 128 256; do RRTK_CF_POOL=$p python bench.py --collision-only --no-cpu 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pool $p', d['ms_per_launch'], d['isotropic_field_kernel']['ms_per_launch'])
"; done
