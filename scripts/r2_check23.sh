python -m pytest tests/test_gpu_rewire.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -8
ncu --set full --clock-control none --import-source on -k regex:collision_cf -s 16 -c 1 -f -o gpurun_out/r2_v4_cfd python bench.py --collision-only --no-cpu > gpurun_out/r2_v4_cfd_bench.log 2>&1
ls -la gpurun_out/r2_v4_cfd.ncu-rep
