mkdir -p gpurun_out
for cfg in "512 0" "448 0" "384 0" "512 512" "384 512" "256 0"; do set -- $cfg
RRTK_PLAN_CAP=$1 RRTK_PLAN_T=$2 timeout 600 python bench.py --informed-only --no-cpu --steps 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cap $1 T $2', round(d['plans_per_s']), 'ms', round(d['ms_per_launch'],1), 'blocks/SM', d['blocks_per_sm'], 'smem', d['smem_bytes_per_block'], 'e2e', round(d['e2e']['value']))"
done
for P in 444 512; do for T in 128 256; do PLANS=$P bash scripts/variants.sh $T main 2>&1 | tail -1; done; done
