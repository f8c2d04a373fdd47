mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pool in 64 128 256; do
RRTK_CF_POOL=$pool timeout 300 python bench.py --collision-only --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pool $pool', 'K1b ms', round(d['ms_per_launch'],4), 'frac', round(d['roofline']['frac'],3), d['same_outputs_as_bit_grid_kernel'], 'K1', round(d['bit_grid_kernel']['ms_per_launch'],4))"
done
bash scripts/sanitize.sh r2 > /dev/null 2>&1; tail -50 gpurun_out/r2_sanitizer.log
