# text summary of one --set full capture, as committed under profiles/: the details-page lines the docs quote plus four raw counters
# usage: scripts/ncu_summary.sh gpurun_out/r2_v3_plan.ncu-rep > profiles/r2_v3_plan_ncu.txt
REP=$1
ncu -i "$REP" --page details 2>/dev/null | grep -E "^  [a-z].*\(|SM Frequency|Elapsed Cycles|Memory Throughput|DRAM Throughput|Duration|L1/TEX Cache Throughput|L2 Cache Throughput|Compute \(SM\) Throughput|Executed Ipc Active|Issue Slots Busy|Mem Busy|L1/TEX Hit Rate|L2 Hit Rate|No Eligible|Eligible Warps|Warp Cycles Per Issued|Avg. Active Threads|Executed Instructions|Registers Per Thread|Dynamic Shared Memory Per Block|Block Limit|Theoretical Occupancy|Achieved Occupancy" | grep -v "Avg. Not Predicated\|^    Total \|Avg. Executed Instructions"
ncu -i "$REP" --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h = rows[0]; u = rows[1]; v = rows[-1]
for a, b, c in zip(h, u, v):
    if a in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'smsp__inst_executed.sum'): print(a, b, c)"
