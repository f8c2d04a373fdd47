mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/variants.sh 0 prev main 2>&1 | tail -2
for v in ilp1 main ilp3 ilp4; do for pool in 128 256; do
L=$PWD/exp_$v.so; [ "$v" = main ] && L=$PWD/rrtplanner_b200/librrtk.so
RRTK_CF_POOL=$pool RRTK_LIB=$L timeout 300 python bench.py --collision-only --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v pool $pool', 'K1b ms', round(d['ms_per_launch'],4), 'frac', round(d['roofline']['frac'],3), d['same_outputs_as_bit_grid_kernel'])"
done; done
