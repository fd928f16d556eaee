"""Static SASS instruction count per source line / per device function for one kernel of libcircall_gpu.so.
usage: python scripts/codelines.py k_wt_warp [ncu_sass.csv]   (with a csv from `ncu --page source --csv --print-source sass`
the dynamic counts are shown too)"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "circall_b200", "libcircall_gpu.so")
kern = sys.argv[1]
dyn = sys.argv[2] if len(sys.argv) > 2 else None
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin") and "sm_100a" in f and "session" in f][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text.") and kern in l][0]
ins, cur, funcs = [], None, []
for l in txt[start + 1:]:
    if (l.startswith(".text.") or l.startswith(".section")) and ins:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l):
        ins.append(cur)
    m = re.match(r"^(\S+):$", l)
    if m and not m.group(1).startswith(".L"):
        funcs.append((len(ins), m.group(1)))
print(len(ins), "instructions =", len(ins) * 16 // 1024, "KB")
funcs.append((len(ins), "<end>"))
dynrows = None
if dyn:
    rows = list(csv.reader(open(dyn)))
    h = rows[1]
    dynrows = [(int(r[h.index("Instructions Executed")] or 0), int(r[h.index("# Samples")] or 0)) for r in rows[2:]]
    tot = sum(x[0] for x in dynrows)
    tots = sum(x[1] for x in dynrows)
for (a, n), (b, _) in zip(funcs, funcs[1:]):
    dem = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    extra = ""
    if dynrows and len(dynrows) == len(ins):
        extra = "  dyn instr %5.1f%%  samples %5.1f%%" % (100 * sum(x[0] for x in dynrows[a:b]) / tot, 100 * sum(x[1] for x in dynrows[a:b]) / tots)
    print("%6d instr  %-60s%s" % (b - a, dem[-60:], extra))
by = collections.Counter(ins)
print("largest source lines (static):")
for k, v in by.most_common(25):
    print("   ", k, v)
