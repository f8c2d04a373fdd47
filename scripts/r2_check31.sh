timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -2
RRTK_PLAN_IMPL=grid timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -m gpu -k "not takes_the_wide_kernel" 2>&1 | tail -2
RRTK_GRID_KEY32=0 RRTK_PLAN_IMPL=grid timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not takes_the_wide_kernel" 2>&1 | tail -2
bash scripts/variants.sh 0 main main 2>&1 | tail -2
RRTK_GRID_KEY32=0 bash scripts/variants.sh 0 main 2>&1 | tail -1
