run() { python bench.py --steps 2 --warmup 3 --plans 2368 --threads $2 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 T=$2', round(d['value']), d['kernel_ms']['plan'], d['roofline']['blocks_per_sm'])"; }
run base 128
RRTK_LIB=$PWD/exp_mb8.so run mb8 128
RRTK_LIB=$PWD/exp_mb10.so run mb10 128
RRTK_LIB=$PWD/exp_mb8.so run mb8 64
run base 64
