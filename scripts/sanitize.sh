# compute-sanitizer passes over the kernels (GPU box): racecheck on the shared-memory kernels, memcheck on the host pipeline
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -x -q -k "test_plan_golden and (blobs or adversarial or wall)" 2>&1 | grep -vE "Host Frame" | tail -4
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_rewire.py -x -q -k "96x96_n500 or 96x96_n600 or several_plans" 2>&1 | grep -vE "Host Frame" | tail -4
timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -x -q -k "sample_stream or rejection_path" 2>&1 | grep -vE "Host Frame" | tail -4
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_replan.py tests/test_gpu_rewire.py -x -q -k "pipelined or carry or inflate or edge_cases or 96x96 or dubins_collision or dubins_sample" 2>&1 | tail -4
