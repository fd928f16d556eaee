"""Attribute an ncu source-page export (SASS rows) to device functions and CUDA source lines.

  ncu -i rep.ncu-rep --page source --csv > src.csv            (one kernel per report section)
  cuobjdump -xelf all libcircall_gpu.so ; nvdisasm -c -g cg_session.sm_100a.cubin > sess.sass
  python scripts/ncu_by_line.py src.csv sess.sass k_wt_warpILb0 [--lines 40]

The cubin must come from the build that was profiled; the two instruction streams are aligned by opcode
(difflib) so small drifts are tolerated and reported.
"""
import csv
import difflib
import re
import sys
from collections import defaultdict


def load_sass(path, kernel_key):
    """-> list of (opcode text, function, file, line) for the .text section whose name holds kernel_key"""
    out = []
    on = False
    func = None
    f = l = None
    for ln in open(path, errors="replace"):
        if ln.startswith(".text."):
            on = kernel_key in ln
            func = "<kernel>"
            continue
        if not on:
            continue
        m = re.match(r"\s*\.type\s+\$\S+\$(\S+),@function", ln)
        if m:
            pending = m.group(1)
            continue
        m = re.match(r"^\$\S+\$(\S+):", ln)
        if m:
            func = m.group(1)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            f, l = m.group(1).split("/")[-1], int(m.group(2))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((m.group(2).strip(), func, f, l))
    return out


def load_ncu(path, kernel_key):
    rows = list(csv.reader(open(path)))
    secs = []
    cur = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            secs.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    return secs


def short(fn):
    m = re.match(r"_ZN3cgw\d+(\w+?)(ILi\d+E)?E", fn)
    return m.group(1) if m else fn


def main():
    src_csv, sass_path, key = sys.argv[1:4]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 30
    want = key.replace("ILb0", "<(bool)0>").replace("ILb1", "<(bool)1>")
    secs = [s for s in load_ncu(src_csv, key) if want in s["name"].replace("k_wt_warp<(bool)", "k_wt_warp<(bool)")]
    if not secs:
        secs = load_ncu(src_csv, key)
    sec = secs[0]
    hdr = sec["hdr"]
    ci = {n: hdr.index(n) for n in hdr}
    sass = load_sass(sass_path, key)
    a = [re.sub(r"\s+", " ", r[ci["Source"]].strip()).split(" ")[0 if not r[ci["Source"]].strip().startswith("@") else 1] for r in sec["rows"]]
    b = [re.sub(r"\s+", " ", s[0]).split(" ")[0 if not s[0].startswith("@") else 1] for s in sass]
    sm = difflib.SequenceMatcher(None, a, b, autojunk=False)
    amap = {}
    for blk in sm.get_matching_blocks():
        for k in range(blk.size):
            amap[blk.a + k] = blk.b + k
    print("kernel:", sec["name"][:90])
    print("ncu rows %d, sass instructions %d, aligned %d" % (len(a), len(b), len(amap)))
    cols = ["Instructions Executed", "Thread Instructions Executed", "# Samples", "stall_long_sb", "stall_wait", "stall_no_inst", "stall_short_sb", "stall_branch_resolving",
            "stall_math", "stall_not_selected", "stall_selected", "L2 Theoretical Sectors Global"]
    cols = [c for c in cols if c in ci]
    byf = defaultdict(lambda: defaultdict(float))
    byl = defaultdict(lambda: defaultdict(float))
    tot = defaultdict(float)
    last = None
    for i, r in enumerate(sec["rows"]):
        j = amap.get(i)
        if j is not None:
            last = sass[j]
        fn, f, l = (last[1], last[2], last[3]) if last else ("?", "?", 0)
        for c in cols:
            try:
                v = float(r[ci[c]])
            except ValueError:
                v = 0.0
            byf[fn][c] += v
            byl[(f, l)][c] += v
            tot[c] += v
    print("\nper function:")
    print("%-28s" % "function" + "".join("%14s" % c[:13] for c in cols))
    for fn, d in sorted(byf.items(), key=lambda kv: -kv[1]["Instructions Executed"]):
        print("%-28s" % short(fn)[:28] + "".join("%14.4g" % d[c] for c in cols))
    print("%-28s" % "TOTAL" + "".join("%14.4g" % tot[c] for c in cols))
    for keycol in ("Instructions Executed", "# Samples"):
        print("\ntop source lines by %s:" % keycol)
        for (f, l), d in sorted(byl.items(), key=lambda kv: -kv[1][keycol])[:nlines]:
            print("%-22s" % ("%s:%s" % (f, l)) + "".join("%14.4g" % d[c] for c in cols))


if __name__ == "__main__":
    main()
