"""Condense an `ncu --page source --csv` SASS dump into basic-block-like runs: consecutive
instructions with the same execution count.  Prints address range, #instr, executions, total
warp instructions, stall samples and a hint of what the run does."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
hdr = rows[1]
ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][0], 16)
ins = []
for r in rows[2:]:
    try:
        ins.append((int(r[0], 16) - base, r[1].strip(), int(r[ci]), int(r[cs]), int(r[ct]),
                    {n: int(r[i] or 0) for i, n in stall_cols}))
    except Exception:
        pass
total = sum(i[2] for i in ins)
runs, cur = [], []
for i in ins:
    if cur and (i[2] != cur[-1][2] or "BAR.SYNC" in cur[-1][1]):
        runs.append(cur)
        cur = []
    cur.append(i)
runs.append(cur)
allst = {}
for i in ins:
    for k, v in i[5].items():
        allst[k] = allst.get(k, 0) + v
print(f"total {total/1e6:.1f}M warp instructions; samples {sum(i[3] for i in ins)}; stalls: " +
      " ".join(f"{k}:{v}" for k, v in sorted(allst.items(), key=lambda kv: -kv[1])[:8]))
acc = 0
for run in runs:
    t = sum(i[2] for i in run)
    acc += t
    if 100 * t / total < minshare:
        continue
    ops = {}
    for i in run:
        op = i[1].split()[0] if not i[1].startswith("@") else i[1].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = " ".join(f"{k}x{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:7])
    lanes = sum(i[4] for i in run) / max(1, t)
    st = {}
    for i in run:
        for k, v in i[5].items():
            st[k] = st.get(k, 0) + v
    stop = " ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{run[0][0]:#07x}-{run[-1][0]:#07x} n={len(run):4d} exec={run[0][2]/1e6:8.2f}M tot={t/1e6:8.1f}M {100*t/total:5.1f}% "
          f"cum={100*acc/total:5.1f}% samp={sum(i[3] for i in run):7d} lanes={lanes:4.1f} | {top} | {stop}")
