ncu --set full --clock-control none --import-source on -k regex:plan_grid_kernel -s 3 -c 1 -f -o gpurun_out/r2_v5_plan python bench.py --steps 1 --warmup 3 --plans 1036 --plan-only > gpurun_out/r2_v5_plan_bench.log 2>&1
ls -la gpurun_out/r2_v5_plan.ncu-rep
