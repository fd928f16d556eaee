"""SASS size and resource usage of every kernel / device function in libcircall_gpu.so (no GPU needed)."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "circall_b200", "libcircall_gpu.so")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for f in sorted(os.listdir(d)):
        if not f.endswith(".cubin") or "sm_100a" not in f:
            continue
        txt = subprocess.run(["nvdisasm", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        name, cnt = None, {}
        for l in txt.split("\n"):
            if l.startswith(".text."):
                name = l.strip()[6:]
            elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
                cnt[name] = cnt.get(name, 0) + 1
        for n, c in sorted(cnt.items(), key=lambda kv: kv[1]):
            dem = subprocess.run(["c++filt", n.rstrip(":")], capture_output=True, text=True).stdout.strip()
            dem = re.sub(r"\(anonymous namespace\)::", "", dem).split("(")[0]
            if pat in dem:
                print("%7.1f KB  %-12s %s" % (c * 16 / 1024, f.split(".")[0], dem[:90]))
