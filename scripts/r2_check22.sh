python -m pytest tests/test_gpu_rewire.py -x -q -m gpu -k "pipelined or several or bad_arguments" 2>&1 | tail -8
python bench.py --dubins-only --no-cpu --steps 3 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
for k in ('dubins_rrtstar', 'euclid_rrtstar_with_rewire'):
    r = d[k]; print(k, round(r['plans_per_s']), 'e2e', round(r['e2e']['value']), r['e2e']['matches_device_arm'], 'trees', round(r['e2e']['trees_mode']['value']), r['e2e']['trees_mode']['matches_device_arm'])
"
python bench.py --collision-only 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfd', d['ms_per_launch'], d['roofline']['frac'], 'shared', d['shared_grid_kernel']['ms_per_launch'], d['shared_grid_kernel']['roofline']['frac'], d['shared_grid_kernel'].get('matches_oracle'), d['shared_grid_kernel']['mean_cells_per_segment'])
"
ncu --set full --clock-control none --import-source on -k regex:collision_cf -s 12 -c 1 -f -o gpurun_out/r2_v4_cfd python bench.py --collision-only --no-cpu > gpurun_out/r2_v4_cfd_bench.log 2>&1
ls -la gpurun_out/r2_v4_cfd.ncu-rep
