"""TEST SCAFFOLDING: ctypes wrapper of tests/emu/libemu.so (host emulation of the device code)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libemu.so")
    srcs = [os.path.join(_HERE, "emu.cpp"),
            os.path.join(_HERE, "..", "..", "circall_b200", "csrc", "cg_core.cuh"),
            os.path.join(_HERE, "..", "..", "circall_b200", "csrc", "cg_pipeline.cuh"),
            os.path.join(_HERE, "..", "..", "circall_b200", "csrc", "cg_thrmem.cuh"),
            os.path.join(_HERE, "..", "..", "circall_b200", "csrc", "cg_pre.cuh"),
            os.path.join(_HERE, "..", "..", "circall_b200", "csrc", "cg_thr.cuh")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", srcs[0], "-o", so])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, u64, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int
        L.emu_index_create.restype = vp
        L.emu_index_create.argtypes = [vp, u64, vp, i32]
        L.emu_index_free.argtypes = [vp]
        L.emu_index_num_kmers.restype = u64
        L.emu_index_num_kmers.argtypes = [vp]
        L.emu_collect.restype = i32
        L.emu_collect.argtypes = [vp, vp, vp, u64, i32, i32, vp, vp, vp, vp, vp, u64, vp, vp, vp, vp, u64]
        L.emu_run.restype = vp
        L.emu_run.argtypes = [vp, vp, vp, vp, vp, vp, u64, i32]
        L.emu_run_free.argtypes = [vp]
        L.emu_run_err.restype = i32
        L.emu_run_err.argtypes = [vp]
        L.emu_run_sizes.argtypes = [vp, vp]
        L.emu_run_fallbacks.argtypes = [vp, vp]
        L.emu_run_tiers.argtypes = [vp, vp]
        L.emu_run_fill.argtypes = [vp] * 12
        L.emu_stdsort.argtypes = [vp, i32]
        L.emu_index_fm.restype = i32
        L.emu_index_fm.argtypes = [vp]
        L.emu_walk_stats.argtypes = [vp, i32]
        _LIB = L
    return _LIB


class EmuIndex:
    def __init__(self, text, sa, k=31):
        text = np.ascontiguousarray(text, np.uint8)
        sa = np.ascontiguousarray(sa, np.int32)
        self._h = _lib().emu_index_create(text.ctypes.data, text.size, sa.ctypes.data, k)
        self.n_kmers = _lib().emu_index_num_kmers(self._h)
        self.fm = _lib().emu_index_fm(self._h)

    def __del__(self):
        try:
            if self._h:
                _lib().emu_index_free(self._h)
                self._h = None
        except Exception:
            pass

    def collect(self, bases, offsets, strict=True, lce_mode=0):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = offsets.size - 1
        cap_i, cap_h = 256 * n + 16, 2048 * n + 16
        found = np.zeros(n, np.uint8)
        fo, ro, ho = (np.zeros(n + 1, np.uint64) for _ in range(3))
        fi, ri = np.zeros((cap_i, 4), np.int32), np.zeros((cap_i, 4), np.int32)
        ht, hp, hf = np.zeros(cap_h, np.uint32), np.zeros(cap_h, np.int32), np.zeros(cap_h, np.uint8)
        rc = _lib().emu_collect(self._h, bases.ctypes.data, offsets.ctypes.data, n, int(strict), lce_mode,
                                found.ctypes.data, fo.ctypes.data, fi.ctypes.data, ro.ctypes.data, ri.ctypes.data,
                                cap_i, ho.ctypes.data, ht.ctypes.data, hp.ctypes.data, hf.ctypes.data, cap_h)
        if rc != 0:
            raise RuntimeError("emu_collect failed: %d" % rc)
        return dict(found=found, fwd_off=fo, fwd_ints=fi[:int(fo[-1])], rc_off=ro, rc_ints=ri[:int(ro[-1])],
                    hits_off=ho, hit_tid=ht[:int(ho[-1])], hit_pos=hp[:int(ho[-1])], hit_fwd=hf[:int(ho[-1])])


def run(tx, bsj, r1, off1, r2, off2, lce_mode=0):
    L = _lib()
    r1 = np.ascontiguousarray(r1, np.uint8); r2 = np.ascontiguousarray(r2, np.uint8)
    off1 = np.ascontiguousarray(off1, np.uint64); off2 = np.ascontiguousarray(off2, np.uint64)
    n = off1.size - 1
    h = L.emu_run(tx._h, bsj._h if bsj else None, r1.ctypes.data, off1.ctypes.data, r2.ctypes.data, off2.ctypes.data,
                  n, lce_mode)
    try:
        if L.emu_run_err(h):
            raise RuntimeError("emu_run error %d" % L.emu_run_err(h))
        sz = np.zeros(5, np.uint64)
        L.emu_run_sizes(h, sz.ctypes.data)
        ne, nj, nc, nt, nm = (int(x) for x in sz)
        out = dict(mate_flags=np.zeros(nm, np.uint8), mate_nhits=np.zeros(nm, np.uint16),
                   exp_pair=np.zeros(ne, np.uint64), exp_code=np.zeros(ne, np.uint8), junc=np.zeros((nj, 6), np.int64),
                   cand=np.zeros(nc, np.uint64), rec_nhits=np.zeros(ne, np.uint32), rec_split=np.zeros(ne, np.uint8),
                   eq_lens=np.zeros(nc, np.uint32), eq_tids=np.zeros(nt, np.uint32), ctr=np.zeros(8, np.uint64))
        fb = np.zeros(2, np.uint64)
        L.emu_run_fallbacks(h, fb.ctypes.data)
        out["fallbacks"] = fb
        ts = np.zeros(3, np.uint64)
        L.emu_run_tiers(h, ts.ctypes.data)
        out["tiers"] = ts        # [wt mates answered by the front kernel's function, wt mates / bsj records answered by the lockstep thread tier]
        L.emu_run_fill(h, *[out[k].ctypes.data for k in ("mate_flags", "mate_nhits", "exp_pair", "exp_code", "junc",
                                                         "cand", "rec_nhits", "rec_split", "eq_lens", "eq_tids", "ctr")])
        return out
    finally:
        L.emu_run_free(h)


def walk_stats(reset=True):
    """(seed-filter look-ups, table look-ups, windows the filter settled) of the lockstep tier's walks since the last reset."""
    a = np.zeros(3, np.uint64)
    _lib().emu_walk_stats(a.ctypes.data, int(reset))
    return tuple(int(x) for x in a)


def stdsort(a):
    a = np.ascontiguousarray(a, np.uint32).copy()
    _lib().emu_stdsort(a.ctypes.data, a.size)
    return a
