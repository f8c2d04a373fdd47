mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/variants.sh 0 prev main 2>&1 | tail -2
for P in 512 1024 2048; do
PLANS=$P bash scripts/variants.sh 128 prev main 2>&1 | tail -2
PLANS=$P bash scripts/variants.sh 256 main b4 2>&1 | tail -2
done
RRTK_LIB=$PWD/exp_clk.so timeout 200 python scripts/phase_clocks.py 444 128 2>&1 | tail -3 | head -2 | cut -c1-330
RRTK_LIB=$PWD/exp_clk.so timeout 200 python scripts/phase_clocks.py 1036 128 2>&1 | tail -3 | head -2 | cut -c1-330
