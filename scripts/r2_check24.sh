python scripts/e2e_trace.py 4096 0 2>&1 | tail -14
for c in 296 444 592 888 1036; do echo chunk $c; python scripts/e2e_trace.py 4096 $c 2>&1 | grep "call ms" | tail -2; done
echo MAXCONN 32; CUDA_DEVICE_MAX_CONNECTIONS=32 python scripts/e2e_trace.py 4096 0 2>&1 | grep "call ms" | tail -2
python -m pytest tests/test_gpu_parity.py tests/test_gpu_rewire.py -x -q -m gpu -k "worlds or pipelined or packed or path_records" 2>&1 | tail -3
