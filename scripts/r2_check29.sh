python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sampl or seed or carry or stream" 2>&1 | tail -2
for v in main st512 main st512; do
  L=$PWD/exp_$v.so; [ "$v" = main ] && L=$PWD/rrtplanner_b200/librrtk.so
  RRTK_LIB=$L timeout 300 python bench.py --steps 5 --warmup 3 --plan-only 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v plans/s', round(d['value']), d['kernel_ms'])"
done
