# round 2, fourth GPU pass: source-level profile of the current K7; under-subscribed launches (strong scaling at 8 GPUs = 512 plans)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:plan_scan_kernel -s 3 -c 1 -f -o gpurun_out/r2_v1_plan \
    python bench.py --steps 1 --warmup 3 --plans 1036 --plan-only > gpurun_out/r2_v1_plan_bench.log 2>&1
for P in 512 1024 2048; do
  PLANS=$P bash scripts/variants.sh 128 main 2>&1 | tail -1
  PLANS=$P bash scripts/variants.sh 256 main b4 2>&1 | tail -2
done
