"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (oracle/README.md).

ctypes binding of oracle/liborc.so, the CPU restatement of Circall's quasi-mapping hot path.
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this module; nothing under circall_b200/ does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force=False):
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"])
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return os.path.join(_HERE, "liborc.so")


def _lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "liborc.so")
    if not os.path.exists(so):
        build()
    L = ctypes.CDLL(so)
    vp, i32, u32, u64, cp = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_char_p
    L.orc_last_error.restype = cp
    L.orc_index_from_text.restype = vp
    L.orc_index_from_text.argtypes = [vp, u64, i32, i32]
    L.orc_index_from_text_sa.restype = vp
    L.orc_index_from_text_sa.argtypes = [vp, u64, i32, vp, i32]
    L.orc_suffix_array.restype = i32
    L.orc_suffix_array.argtypes = [vp, u64, i32, i32, vp]
    L.orc_index_from_records.restype = vp
    L.orc_index_from_records.argtypes = [ctypes.POINTER(cp), ctypes.POINTER(cp), u32, i32, i32]
    L.orc_index_free.argtypes = [vp]
    L.orc_index_text_len.restype = u64
    L.orc_index_text_len.argtypes = [vp]
    L.orc_index_num_tx.restype = u32
    L.orc_index_num_tx.argtypes = [vp]
    L.orc_index_num_kmers.restype = u64
    L.orc_index_num_kmers.argtypes = [vp]
    for f in (L.orc_index_text, L.orc_index_sa, L.orc_index_offsets, L.orc_index_lens):
        f.restype = vp
        f.argtypes = [vp]
    L.orc_index_name.restype = cp
    L.orc_index_name.argtypes = [vp, u32]
    L.orc_index_set_name.argtypes = [vp, u32, cp]
    L.orc_index_tid_at.restype = u32
    L.orc_index_tid_at.argtypes = [vp, u64]
    L.orc_khash_find.restype = i32
    L.orc_khash_find.argtypes = [vp, cp, vp, vp]
    L.orc_extend.restype = i32
    L.orc_extend.argtypes = [vp, i32, ctypes.c_int32, ctypes.c_int32, i32, cp, i32, i32, i32, vp]
    L.orc_lce.restype = i32
    L.orc_lce.argtypes = [vp, i32, ctypes.c_int32, ctypes.c_int32, i32, i32]
    L.orc_collect.restype = i32
    L.orc_collect.argtypes = [vp, cp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32]
    L.orc_run.restype = vp
    L.orc_run.argtypes = [vp, vp, i32, vp, vp, vp, vp, u64, i32, i32]
    L.orc_run_free.argtypes = [vp]
    L.orc_run_counters.argtypes = [vp, i32, vp]
    for f in (L.orc_wt_n_export, L.orc_bsj_n_junc, L.orc_bsj_n_cand, L.orc_bsj_n_rec):
        f.restype = u64
        f.argtypes = [vp]
    L.orc_wt_export.argtypes = [vp, vp, vp]
    L.orc_wt_mates.argtypes = [vp, vp, vp]
    L.orc_bsj_junc.argtypes = [vp, vp]
    L.orc_bsj_cand.argtypes = [vp, vp]
    L.orc_bsj_recs.argtypes = [vp, vp, vp]
    L.orc_eq_sizes.argtypes = [vp, i32, vp]
    L.orc_eq_fill.argtypes = [vp, i32, vp, vp, vp]
    L.orc_merge_fuzzy.restype = i32
    L.orc_merge_fuzzy.argtypes = [i32, i32, vp, vp, vp, i32, vp, vp, vp, i32, u32, u32, vp, i32, vp]
    L.orc_stdsort.argtypes = [vp, i32]
    _LIB = L
    return L


def _err():
    return _lib().orc_last_error().decode()


class Index:
    """[U] RapMapSAIndex restated: '$'-joined text, suffix array, k-mer table, rank."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle index: " + _err())
        self._h = handle
        L = _lib()
        self.text_len = L.orc_index_text_len(handle)
        self.n_tx = L.orc_index_num_tx(handle)
        self.n_kmers = L.orc_index_num_kmers(handle)

    @classmethod
    def from_text(cls, text, k=31, nthreads=None):
        text = np.ascontiguousarray(text, dtype=np.uint8)
        return cls(_lib().orc_index_from_text(text.ctypes.data, text.size, k, nthreads or os.cpu_count() or 1))

    @classmethod
    def from_text_with_sa(cls, text, sa, k=31, nthreads=None):
        text = np.ascontiguousarray(text, dtype=np.uint8)
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        assert sa.size == text.size
        return cls(_lib().orc_index_from_text_sa(text.ctypes.data, text.size, k, sa.ctypes.data, nthreads or os.cpu_count() or 1))

    @classmethod
    def from_records(cls, records, k=31, nthreads=None):
        n = len(records)
        names = (ctypes.c_char_p * n)(*[r[0].encode() for r in records])
        seqs = (ctypes.c_char_p * n)(*[r[1].encode() for r in records])
        return cls(_lib().orc_index_from_records(names, seqs, n, k, nthreads or os.cpu_count() or 1))

    def close(self):
        if self._h:
            _lib().orc_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _arr(self, fn, ctype, n):
        p = fn(self._h)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctype)), shape=(n,))

    @property
    def text(self):
        return self._arr(_lib().orc_index_text, ctypes.c_uint8, self.text_len)

    @property
    def sa(self):
        return self._arr(_lib().orc_index_sa, ctypes.c_int32, self.text_len)

    @property
    def offsets(self):
        return self._arr(_lib().orc_index_offsets, ctypes.c_uint32, self.n_tx)

    @property
    def lens(self):
        return self._arr(_lib().orc_index_lens, ctypes.c_uint32, self.n_tx)

    def name(self, t):
        return _lib().orc_index_name(self._h, t).decode()

    def set_name(self, t, nm):
        _lib().orc_index_set_name(self._h, t, nm.encode())

    def tid_at(self, p):
        return _lib().orc_index_tid_at(self._h, p)

    def khash_find(self, kmer):
        b, e = ctypes.c_int32(), ctypes.c_int32()
        ok = _lib().orc_khash_find(self._h, kmer.encode() if isinstance(kmer, str) else kmer, ctypes.byref(b), ctypes.byref(e))
        return (b.value, e.value) if ok else None

    def extend(self, lb_in, ub_in, start_at, read, start, rc=False, brute=False):
        out = (ctypes.c_int32 * 3)()
        rd = read.encode() if isinstance(read, str) else bytes(read)
        _lib().orc_extend(self._h, int(brute), lb_in, ub_in, start_at, rd, len(rd), start, int(rc), out)
        return tuple(out)

    def lce(self, p1, p2, start_at, stop_at, mode=0):
        return _lib().orc_lce(self._h, mode, p1, p2, start_at, stop_at)

    def collect(self, read, strict=True, lce_mode=0, cap_ints=1024, cap_hits=4096):
        """SACollectorCircall::operator() on one read -> (found, fwd_ints, rc_ints, hits)."""
        rd = read.encode() if isinstance(read, str) else bytes(read)
        f = np.zeros((cap_ints, 4), np.int32)
        r = np.zeros((cap_ints, 4), np.int32)
        ht = np.zeros(cap_hits, np.uint32)
        hp = np.zeros(cap_hits, np.int32)
        hf = np.zeros(cap_hits, np.uint8)
        nf, nr, nh = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rcode = _lib().orc_collect(self._h, rd, len(rd), int(strict), lce_mode, f.ctypes.data, ctypes.byref(nf),
                                   r.ctypes.data, ctypes.byref(nr), ht.ctypes.data, hp.ctypes.data, hf.ctypes.data,
                                   ctypes.byref(nh), cap_ints, cap_hits)
        if rcode < 0:
            raise RuntimeError("oracle collect: capacity too small")
        hits = [(int(ht[i]), int(hp[i]), bool(hf[i])) for i in range(nh.value)]
        return bool(rcode), f[:nf.value].copy(), r[:nr.value].copy(), hits


def _flat(reads):
    """(n, L) uint8 array or (bases, offsets) -> (bases uint8, offsets uint64 of n+1)."""
    if isinstance(reads, tuple):
        b, o = reads
        return np.ascontiguousarray(b, np.uint8), np.ascontiguousarray(o, np.uint64)
    a = np.ascontiguousarray(reads, np.uint8)
    n, L = a.shape
    return a.reshape(-1), (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))


class Run:
    """Result of Circall_wt (mode 0), Circall_bsj (mode 1) or wt->bsj (mode 2) on a batch of pairs."""

    def __init__(self, tx, bsj, mode, r1, r2, nthreads=None, lce_mode=0):
        L = _lib()
        b1, o1 = _flat(r1)
        b2, o2 = _flat(r2)
        self.n = o1.size - 1
        self.mode = mode
        self._h = L.orc_run(tx._h if tx else None, bsj._h if bsj else None, mode, b1.ctypes.data, o1.ctypes.data,
                            b2.ctypes.data, o2.ctypes.data, self.n, nthreads or os.cpu_count() or 1, lce_mode)
        if not self._h:
            raise RuntimeError("oracle run: " + _err())

    def close(self):
        if self._h:
            _lib().orc_run_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counters(self, which):
        out = np.zeros(10, np.uint64)
        _lib().orc_run_counters(self._h, which, out.ctypes.data)
        keys = ["numObserved", "numMapped", "numHits", "upperBound", "readLen", "P", "S", "T", "H", "calls"]
        return dict(zip(keys, (int(x) for x in out)))

    def wt_export(self):
        n = _lib().orc_wt_n_export(self._h)
        pair = np.zeros(n, np.uint64)
        code = np.zeros(n, np.uint8)
        _lib().orc_wt_export(self._h, pair.ctypes.data, code.ctypes.data)
        return pair, code

    def wt_mates(self):
        flags = np.zeros(2 * self.n, np.uint8)
        nh = np.zeros(2 * self.n, np.uint16)
        _lib().orc_wt_mates(self._h, flags.ctypes.data, nh.ctypes.data)
        return flags, nh

    def bsj_junctions(self):
        n = _lib().orc_bsj_n_junc(self._h)
        out = np.zeros((n, 6), np.int64)
        _lib().orc_bsj_junc(self._h, out.ctypes.data)
        return out

    def bsj_candidates(self):
        n = _lib().orc_bsj_n_cand(self._h)
        out = np.zeros(n, np.uint64)
        _lib().orc_bsj_cand(self._h, out.ctypes.data)
        return out

    def bsj_records(self):
        n = _lib().orc_bsj_n_rec(self._h)
        nh = np.zeros(n, np.uint32)
        sp = np.zeros(n, np.uint8)
        _lib().orc_bsj_recs(self._h, nh.ctypes.data, sp.ctypes.data)
        return nh, sp

    def eq_classes(self, which):
        sz = np.zeros(2, np.uint64)
        _lib().orc_eq_sizes(self._h, which, sz.ctypes.data)
        lens = np.zeros(int(sz[0]), np.uint32)
        tids = np.zeros(int(sz[1]), np.uint32)
        cnt = np.zeros(int(sz[0]), np.uint64)
        _lib().orc_eq_fill(self._h, which, lens.ctypes.data, tids.ctypes.data, cnt.ctypes.data)
        out = {}
        p = 0
        for ln, c in zip(lens, cnt):
            out[tuple(int(x) for x in tids[p:p + ln])] = int(c)
            p += int(ln)
        return out


def merge_fuzzy(left, right, left_matches, right_matches, read_len=100, max_num_hits=200):
    """mergeLeftRightHitsFuzzy on two tid-sorted hit lists [(tid, pos, fwd), ...]."""
    def cols(h):
        return (np.array([x[0] for x in h], np.uint32), np.array([x[1] for x in h], np.int32),
                np.array([x[2] for x in h], np.uint8))
    lt, lp, lf = cols(left)
    rt, rp, rf = cols(right)
    cap = len(left) + len(right) + 1
    out = np.zeros((cap, 7), np.int32)
    tm = ctypes.c_int()
    n = _lib().orc_merge_fuzzy(int(left_matches), int(right_matches), lt.ctypes.data, lp.ctypes.data, lf.ctypes.data,
                               len(left), rt.ctypes.data, rp.ctypes.data, rf.ctypes.data, len(right), read_len,
                               max_num_hits, out.ctypes.data, cap, ctypes.byref(tm))
    return out[:n].copy(), bool(tm.value)


def suffix_array(text, algo=0, nthreads=None):
    """Suffix array of a '$'-joined text: algo 0 = induced sorting (the index builder's), 1 = bucketed comparison sort."""
    text = np.ascontiguousarray(text, dtype=np.uint8)
    out = np.zeros(text.size, np.int32)
    if _lib().orc_suffix_array(text.ctypes.data, text.size, algo, nthreads or os.cpu_count() or 1, out.ctypes.data) != 0:
        raise RuntimeError("oracle suffix array: " + _err())
    return out


def stdsort(a):
    """libstdc++ std::sort of packed kmerScores entries by (entry >> 8)."""
    a = np.ascontiguousarray(a, np.uint32).copy()
    _lib().orc_stdsort(a.ctypes.data, a.size)
    return a
