RRTK_PLAN_IMPL=grid timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_fullsize.py -q -m gpu -k "not takes_the_wide_kernel" 2>&1 | tail -15
bash scripts/variants.sh 0 main 2>&1 | tail -1
RRTK_PLAN_IMPL=grid bash scripts/variants.sh 0 main 2>&1 | tail -1
