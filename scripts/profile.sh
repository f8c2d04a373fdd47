# GPU-box profiling pass (B200_PROFILING.md recipe).  usage: scripts/profile.sh <tag> [plans]
# 1. launch list of the bench command (cold-cache, serialised: shares only)
# 2. one --set full capture of the plan kernel (source-level) on a reduced batch
# 3. one --set full capture of the collision kernel (cfg2 microbenchmark)
TAG=${1:-r1}; PLANS=${2:-888}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-dubins > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:plan_scan_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_plan \
    python bench.py --steps 1 --warmup 3 --plans $PLANS --plan-only > gpurun_out/${TAG}_plan_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:collision -s 3 -c 1 -f -o gpurun_out/${TAG}_cc \
    python bench.py --collision-only --no-cpu > gpurun_out/${TAG}_cc_bench.log 2>&1
ls -la gpurun_out
