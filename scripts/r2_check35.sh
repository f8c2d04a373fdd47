LOG=gpurun_out/r2c_sanitizer.log; : > $LOG
run() { echo "== compute-sanitizer --tool $1 :: pytest $2 -k \"$3\"" >> $LOG
  timeout 900 compute-sanitizer --tool $1 --print-limit 10 python -m pytest $2 -x -q -k "$3" 2>&1 | grep -vE "Host Frame|^\s*$" | tail -8 >> $LOG; }
RRTK_PLAN_IMPL=grid run racecheck "tests/test_gpu_parity.py" "test_plan_golden and not wide and not informed and (blobs or adversarial or wall or cfg1)"
run memcheck "tests/test_gpu_parity.py" "cfg3_batch_vs_oracle and star"
run racecheck "tests/test_gpu_parity.py" "cfg3_batch_vs_oracle and star"
cat $LOG
bash scripts/variants.sh 0 main 2>&1 | tail -1
