for impl in scan grid; do
RRTK_PLAN_IMPL=$impl ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:plan_ -s 3 -c 1 python bench.py --steps 1 --warmup 3 --plans 1036 --plan-only 2>&1 | grep -E "plan_|inst_executed|time_duration|issue_active|pipe_|per_inst" | cut -c1-150
done
