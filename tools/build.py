"""Build the drop-in executables: tools/bin/circall_gpu plus the names Circall.sh calls (Circall_wt, Circall_bsj,
Circall_pseudo, TxIndexer) as symlinks to it.  Plain g++; links libcircall_gpu.so (built by circall_b200/build.py) by rpath."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build(force=False):
    lib = os.path.join(ROOT, "circall_b200", "libcircall_gpu.so")
    if not os.path.exists(lib):
        sys.path.insert(0, ROOT)
        from circall_b200 import build as b
        b.build()
    out = os.path.join(HERE, "bin")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "circall_gpu")
    srcs = [os.path.join(HERE, "circall_gpu_main.cpp"), os.path.join(HERE, "cg_host_io.hpp"), os.path.join(ROOT, "include", "circall_gpu.h")]
    if force or not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(s) for s in srcs + [lib]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", srcs[0], "-o", exe, "-L" + os.path.dirname(lib), "-lcircall_gpu",
                               "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath-link,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart", "-lz"])
    # the R stage's joins (SE / PE filters, final table) as a host-only tool
    flt, fsrc = os.path.join(out, "circall_filter"), os.path.join(HERE, "circall_filter.cpp")
    if force or not os.path.exists(flt) or os.path.getmtime(flt) < os.path.getmtime(fsrc):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", fsrc, "-o", flt])
    for name in ("Circall_wt", "Circall_bsj", "Circall_pseudo", "TxIndexer"):
        p = os.path.join(out, name)
        if not os.path.islink(p):
            if os.path.exists(p):
                os.remove(p)
            os.symlink("circall_gpu", p)
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
