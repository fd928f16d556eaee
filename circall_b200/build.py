"""Build libcircall_gpu.so in-tree with nvcc for sm_100a (cross-compiles on a box without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libcircall_gpu.so")
SOURCES = ["cg_api.cu", "cg_index.cu", "cg_session.cu"]
HEADERS = ["cg_common.h", "cg_sortscan.cuh", "cg_core.cuh", "cg_pipeline.cuh", "cg_kernels.cuh", "cg_pack_swar.h", "cg_warp.cuh", "cg_thrmem.cuh", "cg_pre.cuh", "cg_thr.cuh", os.path.join("..", "..", "include", "circall_gpu.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-extended-lambda",
              "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), tag=""):
    """tag/defines: tuning variants (libcircall_gpu<tag>.so built with -D<define>), picked up through CG_LIB_PATH"""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    SO = os.path.join(HERE, "libcircall_gpu%s.so" % tag)
    NVCC_FLAGS = globals()["NVCC_FLAGS"] + ["-D" + d for d in defines]
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(HERE, "build", src.replace(".cu", tag + ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            log = open(o + ".log", "w")
            procs.append((src, subprocess.Popen([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o], stdout=log, stderr=subprocess.STDOUT), log))
    for src, p, log in procs:
        rc = p.wait()
        log.close()
        if rc != 0 or verbose:
            sys.stderr.write(open(log.name).read())
        if rc != 0:
            raise RuntimeError("nvcc failed on " + src)
    if force or procs or _stale(SO, objs):
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO] + objs + ["-lcudart", "-ldl"])
    return SO


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    tag = next((a[6:] for a in sys.argv[1:] if a.startswith("--tag=")), "")
    print(build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, defines=defs, tag=tag))
