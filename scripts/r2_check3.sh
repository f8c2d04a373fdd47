# round 2, third GPU pass: K7 on the round-1 structure + rotating commit warp, adaptive costing slots, fix-up-free division
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash scripts/variants.sh 0 base main fc s2 wf s2wf df 2>&1 | tail -8
for v in s2wf; do
  RRTK_LIB=$PWD/exp_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
done
timeout 120 python scripts/peaks.py 2>&1 | tail -2
