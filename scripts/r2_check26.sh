python -m pytest tests/test_gpu_parity.py tests/test_gpu_rewire.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -3
