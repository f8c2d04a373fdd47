timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for t in 128 64 256; do
  python bench.py --steps 2 --warmup 3 --plans 2048 --threads $t --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('T=$t auto', d['value'], d['kernel_ms'], d['roofline']['blocks_per_sm'], d['roofline']['frac'])"
  RRTK_GRID_SMEM=1 python bench.py --steps 2 --warmup 3 --plans 2048 --threads $t --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('T=$t smem', d['value'], d['kernel_ms'], d['roofline']['blocks_per_sm'], d['roofline']['frac'])"
done
