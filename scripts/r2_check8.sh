mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash scripts/variants.sh 0 main seq 2>&1 | tail -2
RRTK_LIB=$PWD/exp_clk.so timeout 300 python scripts/phase_clocks.py 1036 128 2>&1 | tail -1
for pool in 0 64 128 256; do
RRTK_CF_POOL=$pool timeout 300 python bench.py --collision-only --no-cpu 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pool $pool', 'K1 ms', round(d['ms_per_launch'],4), 'K1b ms', round(d['clearance_field_kernel']['ms_per_launch'],4), d['clearance_field_kernel']['same_outputs_as_bit_grid_kernel'])"
done
