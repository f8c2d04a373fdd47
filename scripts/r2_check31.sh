timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -2
bash scripts/variants.sh 0 main main 2>&1 | tail -2
