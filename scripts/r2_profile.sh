# round 2 profiling pass (B200_PROFILING.md recipe): launch list of the bench command, --set full of K7 (source-level) and of the two collision kernels
TAG=${1:-r2}; PLANS=${2:-1036}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-dubins --no-informed --no-class-api --no-strong > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:plan_(grid|scan)_kernel" -s 3 -c 1 -f -o gpurun_out/${TAG}_plan \
    python bench.py --steps 1 --warmup 3 --plans $PLANS --plan-only > gpurun_out/${TAG}_plan_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:collision_cf -s 3 -c 1 -f -o gpurun_out/${TAG}_cf \
    python bench.py --collision-only --no-cpu > gpurun_out/${TAG}_cf_bench.log 2>&1
# the same kernel on the eight octant fields and on the sixteen half-octant fields: their launches come after the 13 / 26 on the fields before
ncu --set full --clock-control none --import-source on -k regex:collision_cf -s 29 -c 1 -f -o gpurun_out/${TAG}_cfd16 \
    python bench.py --collision-only --no-cpu > gpurun_out/${TAG}_cfd16_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:collision_cf -s 16 -c 1 -f -o gpurun_out/${TAG}_cfd \
    python bench.py --collision-only --no-cpu > gpurun_out/${TAG}_cfd_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:collision_global -s 3 -c 1 -f -o gpurun_out/${TAG}_cc \
    python bench.py --collision-only --no-cpu > gpurun_out/${TAG}_cc_bench.log 2>&1
ls -la gpurun_out | tail -8
