"""Warp instructions and stall samples of plan_grid_kernel per phase (line ranges of csrc/plan_grid.cuh as committed with the
profile): python profiles/region_profile.py sass.csv librrtk.so plan_grid.sm_100a.cubin <mangled-name part>  (see line_profile.py)."""
import csv, re, subprocess, sys, tempfile, os
sass_csv, so, cubin_name, func_pat = sys.argv[1:5]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin_name)], capture_output=True, text=True).stdout
line_of = {}; cur=None; infunc=False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m: infunc = func_pat in m.group(1); continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        # innermost-to-outermost: want the plan_scan.cuh line (outermost)
        f, l = os.path.basename(m.group(1)), int(m.group(2))
        if m.group(3): of, ol = os.path.basename(m.group(3)), int(m.group(4))
        else: of, ol = None, None
        cur = (f, l, of, ol); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv))); hdr = rows[1]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[2][0], 16)
regions = [("setup+prologue", 0, 178), ("owner:pickup", 179, 191), ("owner:scan", 192, 239), ("owner:nearest/dup/walk", 240, 287), ("owner:costing+cands", 288, 378), ("owner:record", 379, 392), ("commit", 393, 528), ("round end", 529, 543), ("goal+out", 544, 700)]
agg = {r[0]: [0, 0] for r in regions}; agg["other"] = [0, 0]; tot=[0,0]
unk = {}
for r in rows[2:]:
    try: off = int(r[0], 16) - base; n, s = int(r[ci]), int(r[cs])
    except Exception: continue
    k = line_of.get(off)
    line = None
    if k:
        if k[0] == "plan_grid.cuh": line = k[1]
        elif k[2] == "plan_grid.cuh": line = k[3]
    name = "other"
    if line is not None:
        for nm, a, b in regions:
            if a <= line <= b: name = nm
    else:
        unk[k] = unk.get(k, 0) + n
    agg[name][0] += n; agg[name][1] += s; tot[0]+=n; tot[1]+=s
for nm, (n, s) in agg.items(): print(f"{nm:28s} instr {n/1e6:8.1f}M {100*n/tot[0]:5.1f}%   samples {s:7d} {100*s/tot[1]:5.1f}%")
for k, n in sorted(unk.items(), key=lambda kv: -kv[1])[:12]: print("   other:", k, f"{n/1e6:.1f}M")
