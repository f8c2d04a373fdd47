RRTK_PLAN_IMPL=grid RRTK_LIB=$PWD/exp_gclk.so python scripts/phase_clocks.py 1036 0 2>&1 | tail -2
RRTK_GRID_K=16 RRTK_PLAN_IMPL=grid timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not takes_the_wide_kernel" 2>&1 | tail -2
RRTK_GRID_K=16 RRTK_PLAN_IMPL=grid bash scripts/variants.sh 0 main main 2>&1 | tail -2
RRTK_GRID_K=16 RRTK_PLAN_IMPL=grid RRTK_LIB=$PWD/exp_gclk.so python scripts/phase_clocks.py 1036 0 2>&1 | tail -2
