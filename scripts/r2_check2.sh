# round 2, second GPU pass: which of the three K7 changes costs / gains what (plans/s, then warp instructions per launch)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash scripts/variants.sh 0 base main oc op oo 2>&1 | tail -6
for v in base main oo; do
  L=$PWD/exp_$v.so; [ "$v" = main ] && L=$PWD/rrtplanner_b200/librrtk.so
  RRTK_LIB=$L timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,smsp__warps_active.avg.per_cycle_active \
     --clock-control none -k regex:plan_scan -s 3 -c 1 --csv --log-file gpurun_out/r2_inst_$v.csv \
     python bench.py --steps 1 --warmup 3 --plans 1036 --plan-only > /dev/null 2>&1
  echo "== $v"; grep -v "^==" gpurun_out/r2_inst_$v.csv | cut -d, -f5,13- | tail -5
done
