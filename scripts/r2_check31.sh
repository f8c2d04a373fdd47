RRTK_PLAN_IMPL=grid timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu -k "not takes_the_wide_kernel" 2>&1 | tail -3
RRTK_PLAN_IMPL=grid bash scripts/variants.sh 0 main main 2>&1 | tail -2
for by in 1 3 4; do RRTK_GRID_BSY=$by RRTK_PLAN_IMPL=grid bash scripts/variants.sh 0 main 2>&1 | tail -1; done
RRTK_GRID_BSX=4 RRTK_GRID_BSY=3 RRTK_PLAN_IMPL=grid bash scripts/variants.sh 0 main 2>&1 | tail -1
RRTK_GRID_BSX=6 RRTK_GRID_BSY=1 RRTK_PLAN_IMPL=grid bash scripts/variants.sh 0 main 2>&1 | tail -1
