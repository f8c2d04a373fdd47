mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/variants.sh 0 prev main 2>&1 | tail -2
bash scripts/variants.sh 0 prev main 2>&1 | tail -2
