mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/variants.sh 0 prev main 2>&1 | tail -2
RRTK_PLAN_CAP=208 bash scripts/variants.sh 0 main b8 2>&1 | tail -2
RRTK_PLAN_CAP=192 bash scripts/variants.sh 0 b8 2>&1 | tail -1
RRTK_PLAN_CAP=208 RRTK_LIB=$PWD/exp_b8.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
