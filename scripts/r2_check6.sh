mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash scripts/variants.sh 0 main b8 2>&1 | tail -2
RRTK_LIB=$PWD/exp_b8.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
RRTK_LIB=$PWD/exp_clk.so timeout 300 python scripts/phase_clocks.py 1036 128 2>&1 | tail -1
RRTK_LIB=$PWD/exp_clk8.so timeout 300 python scripts/phase_clocks.py 1184 128 2>&1 | tail -1
RRTK_LIB=$PWD/exp_clk.so timeout 300 python scripts/phase_clocks.py 444 256 2>&1 | tail -1
RRTK_LIB=$PWD/exp_clk.so timeout 300 python scripts/phase_clocks.py 512 128 2>&1 | tail -1
