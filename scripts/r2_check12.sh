mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "free_cells or packed" 2>&1 | tail -3
RRTK_LIB=$PWD/exp_clk.so timeout 300 python scripts/phase_clocks.py 1036 128 2>&1 | tail -3 | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 --no-collision --no-dubins --no-strong --no-e2e > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -3 gpurun_out/r2_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench2.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('informed_bench'))[:1500]); print(json.dumps(d.get('class_api_bench'))[:800])
PY
