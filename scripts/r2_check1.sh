# round 2, first GPU pass: parity suite on the new K7, then base vs new device-arm plans/s, phase clocks
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
bash scripts/variants.sh 0 base main 2>&1 | tail -4
RRTK_LIB=$PWD/exp_clk.so timeout 300 python scripts/phase_clocks.py 1036 2>&1 | tail -3
bash scripts/variants.sh 256 main 2>&1 | tail -2
