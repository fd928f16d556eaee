"""The C-ABI library loads on a CPU-only box, exports every symbol include/circall_gpu.h declares, and every
compute entry point fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

import circall_b200 as cg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "circall_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from circall_b200 import build
    build.build()
    L = cg.lib()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
        getattr(L, n)


def test_header_is_plain_c():
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write('#include "circall_gpu.h"\nint main(void){cg_opts o; cg_counters c; (void)o; (void)c; return CG_MAX_READ_LEN == 256 ? 0 : 1;}\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", c, "-o", os.path.join(d, "t.o")])


@pytest.mark.skipif(cg.device_count() > 0, reason="checks the no-GPU behaviour")
def test_compute_fails_loudly_without_gpu():
    text = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT$", np.uint8)
    with pytest.raises(cg.CircallError) as e:
        cg.Index.build(text)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(cg.CircallError):
        cg.Index.build_fasta("/nonexistent.fa")
    with pytest.raises(cg.CircallError):
        cg.bench_gather(1 << 24)


def test_argument_validation():
    L = cg.lib()
    h = ctypes.c_void_p()
    assert L.cg_index_build(0, None, 10, 31, ctypes.byref(h)) == -1
    t = np.frombuffer(b"ACGT", np.uint8)
    assert L.cg_index_build(0, t.ctypes.data, 4, 31, ctypes.byref(h)) == -1          # no trailing '$'
    t = np.frombuffer(b"ACGT$", np.uint8)
    assert L.cg_index_build(0, t.ctypes.data, 5, 30, ctypes.byref(h)) == -1          # k must be odd
    assert b"odd" in L.cg_last_error()
    assert L.cg_session_create(None, None, None, ctypes.byref(h)) == -1


def test_product_does_not_touch_the_oracle():
    """Nothing under circall_b200/ may import, link or execute oracle/ (or the test emulation)."""
    for dp, dn, fn in os.walk(os.path.join(ROOT, "circall_b200")):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                s = open(os.path.join(dp, f), errors="replace").read()
                assert "liborc" not in s and "import oracle" not in s and "oracle/" not in s.replace("the oracle", ""), f
                assert "libemu" not in s
