"""The device code's host emulation (tests/emu: cg_core.cuh / cg_pre.cuh / cg_thr.cuh compiled for the host) under AddressSanitizer and
UndefinedBehaviorSanitizer: the per-thread scratch indexing, the shifts of the walk's dead-window masks, the lcp-byte extension and the
interval hand-over are exercised on a pipeline's worth of reads with every access checked.  (compute-sanitizer is not available on the
GPU pool; this is the bounds check the thread tiers get.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, ctypes
sys.path.insert(0, %(root)r)
import numpy as np
import oracle, synth
from tests import emu
emu._LIB = None
emu.build = lambda force=False: %(so)r
T = synth.Txome(11, 160, p_antisense=0.05, p_repeat=0.08, p_polya=0.02)
otx, obsj = oracle.Index.from_text(T.tx_text), oracle.Index.from_text(T.bsj_text)
etx, ebsj = emu.EmuIndex(T.tx_text, otx.sa), emu.EmuIndex(T.bsj_text, obsj.sa)
r1, r2 = T.reads(5, 4000, L=100, circ_frac=0.1, n_rate=0.001)
(b1, o1), (b2, o2) = oracle._flat(r1), oracle._flat(r2)
E = emu.run(etx, ebsj, b1, o1, b2, o2, 0)
R = oracle.Run(otx, obsj, 2, r1, r2)
p, c = R.wt_export()
assert np.array_equal(E["exp_pair"], p) and np.array_equal(E["exp_code"], c)
assert sorted(map(tuple, R.bsj_junctions().tolist())) == sorted(map(tuple, E["junc"].tolist()))
print("sanitized run ok", len(p), int(E["tiers"][1]), int(E["tiers"][2]))
"""


def test_thread_tiers_under_asan_ubsan(tmp_path):
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    ubsan = subprocess.run(["gcc", "-print-file-name=libubsan.so"], capture_output=True, text=True).stdout.strip()
    if not (os.path.isabs(asan) and os.path.exists(asan) and os.path.isabs(ubsan) and os.path.exists(ubsan)):
        pytest.skip("no sanitizer runtime in this image")
    so = str(tmp_path / "libemu_san.so")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-fno-omit-frame-pointer", "-x", "c++", os.path.join(ROOT, "tests", "emu", "emu.cpp"), "-o", so])
    env = dict(os.environ, LD_PRELOAD=asan + ":" + ubsan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    p = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "so": so}], capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0 and "sanitized run ok" in p.stdout, (p.stdout[-2000:], p.stderr[-4000:])
