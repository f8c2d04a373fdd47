"""Join an `ncu --page source --csv` SASS dump with nvdisasm line info to get executed warp
instructions and stall samples per CUDA source line.

    ncu -i rep.ncu-rep --page source --csv > sass.csv
    python profiles/line_profile.py sass.csv rrtplanner_b200/librrtk.so plan.sm_100a.cubin 'plan_kernelILi1ELb1E' [top]
"""
import csv
import re
import subprocess
import sys
import tempfile
import os

sass_csv, so, cubin_name, func_pat = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
MAIN = sys.argv[6] if len(sys.argv) > 6 else "plan_scan.cuh"
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin_name)], capture_output=True, text=True).stdout
# walk the disassembly of the wanted function: remember the current source line per instruction offset
line_of = {}
cur, infunc = None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        infunc = func_pat in m.group(1)
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), (os.path.basename(m.group(3)), int(m.group(4))) if m.group(3) else None)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[2][0], 16)
agg = {}
total = 0
for r in rows[2:]:
    try:
        off = int(r[0], 16) - base
        n, s = int(r[ci]), int(r[cs])
    except Exception:
        continue
    key = line_of.get(off)
    # attribute inlined helpers to the plan.cu line that called them
    if key and key[2] and key[0] != MAIN:
        key = (key[2][0], key[2][1], None, "via " + key[0])
    agg.setdefault(key[:2] if key else None, [0, 0])
    agg[key[:2] if key else None][0] += n
    agg[key[:2] if key else None][1] += s
    total += n
src = {}
print(f"total warp instructions {total/1e6:.1f} M")
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if k:
        path = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", k[0])
        if path not in src and os.path.exists(path):
            src[path] = open(path).read().splitlines()
        if path in src and k[1] <= len(src[path]):
            text = src[path][k[1] - 1].strip()[:100]
    print(f"{n/1e6:9.1f}M {100*n/total:5.1f}%  samples {s:6d}  {k}  {text}")
