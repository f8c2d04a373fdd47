mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash scripts/variants.sh 0 main seq 2>&1 | tail -2
RRTK_LIB=$PWD/exp_clk.so timeout 300 python scripts/phase_clocks.py 1036 128 2>&1 | tail -1
RRTK_PLAN_K=16 bash scripts/variants.sh 128 k16 2>&1 | tail -1
RRTK_PLAN_K=16 bash scripts/variants.sh 256 k16 2>&1 | tail -1
RRTK_PLAN_K=8 bash scripts/variants.sh 256 k16 2>&1 | tail -1
RRTK_PLAN_K=16 RRTK_LIB=$PWD/exp_k16.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
