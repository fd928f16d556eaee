"""circall_b200 — B200 (sm_100a) implementation of Circall's quasi-mapping hot path.

Python host side over the C ABI of ``libcircall_gpu.so`` (``include/circall_gpu.h``).  The names
follow the reference (paths relative to datngu/Circall ``Circall_v1.0.1/``):

* :class:`Index`    — the quasi-index a ``TxIndexer`` run produces (``src/TxIndexer.cpp:202-227``) and
  ``SailfishIndex::load`` reads (``include/ReadExperiment.hpp:66-67``);
* :func:`collect`   — ``SACollectorCircall::operator()`` (``include/SACollectorCircall.hpp:24-31``);
* :func:`merge_pairs` — ``mergeLeftRightHitsFuzzy`` as ``src/Circall_pseudo.cpp:271-280`` calls it;
* :class:`Session`  — the per-pair bodies of ``Circall_wt`` (``src/Circall_wt.cpp:237-660``) and
  ``Circall_bsj`` (``src/Circall_bsj.cpp:225-505``), fused.

There is no CPU path: importing works anywhere (so the C ABI can be inspected), but every compute
entry point raises :class:`CircallError` when the CUDA library or a device is missing.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CG_LIB_PATH") or os.path.join(_HERE, "libcircall_gpu.so")   # CG_LIB_PATH: tuning builds
_LIB = None

CG_MAX_READ_LEN = 256
STATUS = {0: "CG_OK", -1: "CG_ERR_ARG", -2: "CG_ERR_CUDA", -3: "CG_ERR_IO", -4: "CG_ERR_BASE", -5: "CG_ERR_LENGTH",
          -6: "CG_ERR_CAPACITY", -7: "CG_ERR_STATE"}


class CircallError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (STATUS.get(code, str(code)), msg))
        self.code = code


class _IndexInfo(ctypes.Structure):
    _fields_ = [("text_len", ctypes.c_uint64), ("n_tx", ctypes.c_uint32), ("n_kmers", ctypes.c_uint64),
                ("hash_capacity", ctypes.c_uint64), ("k", ctypes.c_int), ("device", ctypes.c_int),
                ("device_bytes", ctypes.c_uint64), ("build_seconds", ctypes.c_double)]


class _CollectResult(ctypes.Structure):
    _fields_ = [("n_reads", ctypes.c_uint64), ("found", ctypes.c_void_p),
                ("fwd_off", ctypes.c_void_p), ("fwd_ints", ctypes.c_void_p),
                ("rc_off", ctypes.c_void_p), ("rc_ints", ctypes.c_void_p),
                ("hits_off", ctypes.c_void_p), ("hit_tid", ctypes.c_void_p), ("hit_pos", ctypes.c_void_p),
                ("hit_fwd", ctypes.c_void_p)]


class _Opts(ctypes.Structure):
    _fields_ = [("lce_mode", ctypes.c_int), ("max_batch_pairs", ctypes.c_uint32), ("stage", ctypes.c_int),
                ("keep_mate_detail", ctypes.c_int), ("dump_intervals", ctypes.c_int)]


class _BatchResult(ctypes.Structure):
    _fields_ = [("n_pairs", ctypes.c_uint64),
                ("n_export", ctypes.c_uint64), ("exp_pair", ctypes.c_void_p), ("exp_code", ctypes.c_void_p),
                ("n_junc", ctypes.c_uint64), ("junc", ctypes.c_void_p),
                ("n_cand", ctypes.c_uint64), ("cand_rec", ctypes.c_void_p),
                ("n_eq", ctypes.c_uint64), ("eq_off", ctypes.c_void_p), ("eq_tids", ctypes.c_void_p),
                ("mate_flags", ctypes.c_void_p), ("mate_nhits", ctypes.c_void_p),
                ("rec_nhits", ctypes.c_void_p), ("rec_split", ctypes.c_void_p),
                ("device_ms", ctypes.c_double), ("kernel_ms", ctypes.c_double * 4),
                ("n_fallback_wt", ctypes.c_uint64), ("n_fallback_bsj", ctypes.c_uint64),
                ("n_tier2_wt", ctypes.c_uint64), ("n_tier2_bsj", ctypes.c_uint64),
                ("wt_iv_head", ctypes.c_void_p), ("wt_iv_off", ctypes.c_void_p), ("wt_iv", ctypes.c_void_p), ("n_wt_iv", ctypes.c_uint64),
                ("bsj_iv_head", ctypes.c_void_p), ("bsj_iv_off", ctypes.c_void_p), ("bsj_iv", ctypes.c_void_p), ("n_bsj_iv", ctypes.c_uint64)]


class _Counters(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("wt_num_observed", "wt_num_mapped", "wt_num_hits", "wt_upper_bound",
                                               "bsj_num_observed", "bsj_num_mapped", "bsj_num_hits", "bsj_upper_bound")] + \
               [("wt_read_len", ctypes.c_uint32), ("bsj_read_len", ctypes.c_uint32)]


JUNC_DTYPE = np.dtype([("rec", np.uint32), ("dir", np.uint32), ("query_pos", np.uint32), ("len", np.uint32),
                       ("hit_pos", np.int32), ("tid", np.uint32)])


def lib():
    """The loaded C-ABI library; raises when it has not been built (there is no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise CircallError(-2, "libcircall_gpu.so is not built (run `python -m circall_b200.build`); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, cp, i32, u32, u64 = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64
    pp = ctypes.POINTER(vp)
    L.cg_last_error.restype = cp
    L.cg_device_count.argtypes = [ctypes.POINTER(i32)]
    L.cg_index_build.argtypes = [i32, vp, u64, i32, pp]
    L.cg_index_build_fasta.argtypes = [i32, cp, i32, pp]
    L.cg_index_save.argtypes = [vp, cp]
    L.cg_index_open.argtypes = [cp, i32, pp]
    L.cg_index_close.argtypes = [vp]
    L.cg_index_close.restype = None
    L.cg_index_get_info.argtypes = [vp, ctypes.POINTER(_IndexInfo)]
    L.cg_index_dir_info.argtypes = [ctypes.c_char_p, ctypes.POINTER(_IndexInfo), ctypes.POINTER(ctypes.c_int)]
    L.cg_debug_sort_pairs.argtypes = [ctypes.c_int, vp, vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
    L.cg_debug_scan.argtypes = [ctypes.c_int, vp, ctypes.c_uint64, ctypes.c_int, vp, vp]
    L.cg_index_tx_name.argtypes = [vp, u32]
    L.cg_index_tx_name.restype = cp
    L.cg_index_set_tx_name.argtypes = [vp, u32, cp]
    L.cg_index_tx_len.argtypes = [vp, u32]
    L.cg_index_tx_len.restype = u32
    L.cg_index_tx_offset.argtypes = [vp, u32]
    L.cg_index_tx_offset.restype = u32
    L.cg_index_copy_sa.argtypes = [vp, vp]
    L.cg_collect.argtypes = [vp, vp, vp, u64, i32, i32, ctypes.POINTER(_CollectResult)]
    L.cg_collect_free.restype = None
    L.cg_merge_pairs.argtypes = [i32, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, u32, u32, vp, vp, u64, vp]
    L.cg_session_create.argtypes = [vp, vp, ctypes.POINTER(_Opts), pp]
    L.cg_session_destroy.argtypes = [vp]
    L.cg_session_destroy.restype = None
    L.cg_session_submit.argtypes = [vp, vp, vp, vp, vp, u64]
    L.cg_session_submit_device.argtypes = [vp, vp, vp, vp, vp, u64, u32]
    L.cg_session_submit_fixed.argtypes = [vp, vp, vp, u64, u32]
    L.cg_session_fetch.argtypes = [vp, i32, ctypes.POINTER(_BatchResult)]
    L.cg_session_counters.argtypes = [vp, ctypes.POINTER(_Counters), vp, u64]
    L.cg_session_accumulator.argtypes = [vp, pp, ctypes.POINTER(u64)]
    L.cg_sessions_allreduce.argtypes = [ctypes.POINTER(vp), i32]
    L.cg_session_launches.argtypes = [vp]
    L.cg_session_launches.restype = u64
    L.cg_device_alloc.argtypes = [i32, u64, pp]
    L.cg_device_free.argtypes = [i32, vp]
    L.cg_device_upload.argtypes = [i32, vp, vp, u64]
    L.cg_device_download.argtypes = [i32, vp, vp, u64]
    L.cg_host_alloc.argtypes = [u64, pp]
    L.cg_host_free.argtypes = [vp]
    L.cg_device_flush_l2.argtypes = [i32]
    L.cg_bench_gather.argtypes = [i32, u64, u32, i32, ctypes.POINTER(ctypes.c_double)]
    L.cg_bench_gather_mode.argtypes = [i32, u64, u32, i32, i32, ctypes.POINTER(ctypes.c_double)]
    L.cg_pack_reads.argtypes = [i32, vp, vp, vp, vp, u64, i32, i32, i32, vp, vp, vp, ctypes.POINTER(ctypes.c_double)]
    _LIB = L
    return L


def _check(rc):
    if rc != 0:
        raise CircallError(rc, lib().cg_last_error().decode(errors="replace"))


def device_count():
    n = ctypes.c_int()
    _check(lib().cg_device_count(ctypes.byref(n)))
    return n.value


def _view(ptr, dtype, n):
    """Copy n items of dtype out of library-owned host memory."""
    dtype = np.dtype(dtype)
    if n == 0 or not ptr:
        return np.zeros(0, dtype)
    buf = (ctypes.c_char * (int(n) * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(n)).copy()


def flatten_reads(reads):
    """(n, L) uint8 array, list of str/bytes, or (bases, offsets) -> (bases uint8, offsets uint64 of n+1)."""
    if isinstance(reads, tuple):
        b, o = reads
        return np.ascontiguousarray(b, np.uint8).reshape(-1), np.ascontiguousarray(o, np.uint64)
    if isinstance(reads, (list,)):
        bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
        off = np.zeros(len(bs) + 1, np.uint64)
        if bs:
            off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
        return np.frombuffer(b"".join(bs), np.uint8).copy() if bs else np.zeros(0, np.uint8), off
    a = np.ascontiguousarray(reads, np.uint8)
    n, L = a.shape
    return a.reshape(-1), np.arange(n + 1, dtype=np.uint64) * np.uint64(L)


def debug_sort_pairs(keys, vals, begin_bit=0, end_bit=64, device=0):
    """The index builder's radix sort on host arrays: returns (keys, vals) sorted stably by key bits [begin_bit, end_bit)."""
    k = np.ascontiguousarray(keys, np.uint64).copy()
    v = np.ascontiguousarray(vals, np.uint32).copy()
    _check(lib().cg_debug_sort_pairs(device, k.ctypes.data, v.ctypes.data, k.size, begin_bit, end_bit))
    return k, v


def debug_scan(data, mode, device=0):
    """mode 0: exclusive sum, 1: inclusive max (uint32, wraps like the device); 2: indices of the non-zero flags."""
    d = np.ascontiguousarray(data, np.uint32).copy()
    if mode == 2:
        out = np.zeros(d.size, np.uint32)
        n = ctypes.c_uint64(0)
        _check(lib().cg_debug_scan(device, d.ctypes.data, d.size, 2, out.ctypes.data, ctypes.byref(n)))
        return out[:n.value]
    _check(lib().cg_debug_scan(device, d.ctypes.data, d.size, mode, None, None))
    return d


class Index:
    """Device-resident quasi-index: 2-bit text blocks, suffix array, k-mer table, transcript table."""

    def __init__(self, handle):
        self._h = handle
        info = _IndexInfo()
        _check(lib().cg_index_get_info(self._h, ctypes.byref(info)))
        self.text_len, self.n_tx, self.n_kmers, self.k = info.text_len, info.n_tx, info.n_kmers, info.k
        self.hash_capacity, self.device, self.device_bytes = info.hash_capacity, info.device, info.device_bytes
        self.build_seconds = info.build_seconds

    @classmethod
    def build(cls, text, k=31, device=0):
        """text: the '$'-joined, already normalised transcripts (uint8 array / bytes)."""
        t = np.ascontiguousarray(np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray)) else text, np.uint8)
        h = ctypes.c_void_p()
        _check(lib().cg_index_build(device, t.ctypes.data, t.size, k, ctypes.byref(h)))
        return cls(h)

    @classmethod
    def build_fasta(cls, path, k=31, device=0):
        h = ctypes.c_void_p()
        _check(lib().cg_index_build_fasta(device, os.fsencode(path), k, ctypes.byref(h)))
        return cls(h)

    @classmethod
    def open(cls, directory, device=0):
        h = ctypes.c_void_p()
        _check(lib().cg_index_open(os.fsencode(directory), device, ctypes.byref(h)))
        return cls(h)

    def save(self, directory):
        _check(lib().cg_index_save(self._h, os.fsencode(directory)))

    @staticmethod
    def dir_info(directory):
        """What an index directory holds, read on the host (no device): dict(text_len, n_tx, k, n_kmers, hash_capacity, layout);
        layout 1 = the reference's file set only (txpInfo.bin ...), 2 = with this library's cache beside it."""
        info, lay = _IndexInfo(), ctypes.c_int(0)
        _check(lib().cg_index_dir_info(os.fsencode(directory), ctypes.byref(info), ctypes.byref(lay)))
        return dict(text_len=info.text_len, n_tx=info.n_tx, k=info.k, n_kmers=info.n_kmers, hash_capacity=info.hash_capacity, layout=lay.value)

    def close(self):
        if self._h:
            lib().cg_index_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tx_name(self, t):
        return lib().cg_index_tx_name(self._h, t).decode()

    def set_tx_name(self, t, name):
        _check(lib().cg_index_set_tx_name(self._h, t, name.encode()))

    def tx_len(self, t):
        return lib().cg_index_tx_len(self._h, t)

    def tx_offset(self, t):
        return lib().cg_index_tx_offset(self._h, t)

    def suffix_array(self):
        sa = np.empty(self.text_len, np.int32)
        _check(lib().cg_index_copy_sa(self._h, sa.ctypes.data))
        return sa


def collect(index, reads, strict=True, lce_mode=0):
    """SACollectorCircall::operator() on every read: post-vote intervals and the hit list per read."""
    b, o = flatten_reads(reads)
    res = _CollectResult()
    _check(lib().cg_collect(index._h, b.ctypes.data, o.ctypes.data, o.size - 1, int(strict), lce_mode, ctypes.byref(res)))
    n = o.size - 1
    fo, ro, ho = _view(res.fwd_off, np.uint64, n + 1), _view(res.rc_off, np.uint64, n + 1), _view(res.hits_off, np.uint64, n + 1)
    out = dict(found=_view(res.found, np.uint8, n), fwd_off=fo, fwd_ints=_view(res.fwd_ints, np.int32, 4 * int(fo[-1])).reshape(-1, 4),
               rc_off=ro, rc_ints=_view(res.rc_ints, np.int32, 4 * int(ro[-1])).reshape(-1, 4), hits_off=ho,
               hit_tid=_view(res.hit_tid, np.uint32, int(ho[-1])), hit_pos=_view(res.hit_pos, np.int32, int(ho[-1])),
               hit_fwd=_view(res.hit_fwd, np.uint8, int(ho[-1])))
    lib().cg_collect_free()
    return out


def merge_pairs(left, right, left_seeded, right_seeded, read_len=100, max_num_hits=200, device=0):
    """mergeLeftRightHitsFuzzy for a batch of pairs.

    left/right: per pair a list of (tid, pos, fwd) sorted by tid.  Returns (joint_off, joint[n,7], too_many).
    """
    def cols(lists):
        off = np.zeros(len(lists) + 1, np.uint64)
        off[1:] = np.cumsum([len(x) for x in lists], dtype=np.uint64) if lists else 0
        flat = [h for x in lists for h in x]
        return (off, np.array([h[0] for h in flat], np.uint32), np.array([h[1] for h in flat], np.int32),
                np.array([h[2] for h in flat], np.uint8))
    n = len(left)
    lo, lt, lp, lf = cols(left)
    ro, rt, rp, rf = cols(right)
    ls = np.ascontiguousarray(left_seeded, np.uint8)
    rs = np.ascontiguousarray(right_seeded, np.uint8)
    cap = int(lo[-1] + ro[-1]) + n + 1
    joff = np.zeros(n + 1, np.uint64)
    joint = np.zeros((cap, 7), np.int32)
    tm = np.zeros(max(n, 1), np.uint8)
    _check(lib().cg_merge_pairs(device, n, lo.ctypes.data, lt.ctypes.data, lp.ctypes.data, lf.ctypes.data, ls.ctypes.data,
                                ro.ctypes.data, rt.ctypes.data, rp.ctypes.data, rf.ctypes.data, rs.ctypes.data, read_len,
                                max_num_hits, joff.ctypes.data, joint.ctypes.data, cap, tm.ctypes.data))
    return joff, joint[:int(joff[-1])].copy(), tm[:n].astype(bool)


class BatchResult:
    """Host copy of one fetched batch (see cg_batch_result in include/circall_gpu.h)."""

    def __init__(self, r, want_lists, detail):
        self.n_pairs, self.n_export, self.n_junc, self.n_eq = r.n_pairs, r.n_export, r.n_junc, r.n_eq
        self.device_ms = r.device_ms
        self.kernel_ms = tuple(r.kernel_ms)      # pack, wt, commit+compaction, bsj
        self.n_fallback_wt, self.n_fallback_bsj = r.n_fallback_wt, r.n_fallback_bsj
        self.n_tier2_wt, self.n_tier2_bsj = r.n_tier2_wt, r.n_tier2_bsj
        self.n_cand = r.n_cand
        # bytes cg_session_fetch copied device -> host for this batch: the batch header, and with the lists every array it exposes
        self.d2h_bytes = 64
        if r.exp_pair:
            n_eq_tids = int(ctypes.cast(r.eq_off, ctypes.POINTER(ctypes.c_uint64))[r.n_eq]) if r.eq_off else 0
            self.d2h_bytes += r.n_export * (4 + 1 + 4 + 1) + r.n_junc * JUNC_DTYPE.itemsize + 4 * r.n_cand + 16 * r.n_eq + 4 * n_eq_tids
            if r.mate_flags:
                self.d2h_bytes += 2 * r.n_pairs * 3
        if not want_lists:
            return
        self.exp_pair = _view(r.exp_pair, np.uint32, r.n_export)
        self.exp_code = _view(r.exp_code, np.uint8, r.n_export)
        self.junc = _view(r.junc, JUNC_DTYPE, r.n_junc)
        self.n_cand = r.n_cand
        self.cand_rec = _view(r.cand_rec, np.uint32, r.n_cand)
        self.eq_off = _view(r.eq_off, np.uint64, r.n_eq + 1 if r.eq_off else 0)
        self.eq_tids = _view(r.eq_tids, np.uint32, int(self.eq_off[-1]) if self.eq_off.size else 0)
        self.rec_nhits = _view(r.rec_nhits, np.uint32, r.n_export)
        self.rec_split = _view(r.rec_split, np.uint8, r.n_export)
        if detail:
            self.mate_flags = _view(r.mate_flags, np.uint8, 2 * r.n_pairs)
            self.mate_nhits = _view(r.mate_nhits, np.uint16, 2 * r.n_pairs)
        # cg_opts::dump_intervals: (head, off, intervals[n, 4]) per Circall_wt mate / per Circall_bsj record
        self.wt_iv = self.bsj_iv = None
        if r.wt_iv_head:
            self.wt_iv = (_view(r.wt_iv_head, np.uint32, 2 * r.n_pairs), _view(r.wt_iv_off, np.uint32, 2 * r.n_pairs),
                          _view(r.wt_iv, np.int32, 4 * r.n_wt_iv).reshape(-1, 4))
        if r.bsj_iv_head:
            self.bsj_iv = (_view(r.bsj_iv_head, np.uint32, r.n_export), _view(r.bsj_iv_off, np.uint32, r.n_export),
                           _view(r.bsj_iv, np.int32, 4 * r.n_bsj_iv).reshape(-1, 4))

    @staticmethod
    def intervals_of(iv, i):
        """(found, tier, fwd[nF, 4], rc[nR, 4]) of item i of a dump_intervals triple; tier 0 = no tier recorded it."""
        head, off, pool = iv
        h, o = int(head[i]), int(off[i])
        nF, nR, tier, found = h & 0xfff, (h >> 12) & 0xfff, (h >> 24) & 0xf, bool((h >> 28) & 1)
        return found, tier, pool[o:o + nF], pool[o + nF:o + nF + nR]

    def eq_classes(self):
        """{sorted tid tuple -> count} of the multi-transcript classes of this batch."""
        out = {}
        for c in range(self.eq_off.size - 1):
            key = tuple(int(x) for x in self.eq_tids[int(self.eq_off[c]):int(self.eq_off[c + 1])])
            out[key] = out.get(key, 0) + 1
        return out


class Session:
    """Circall_wt + Circall_bsj on batches of read pairs.  stage: 0 fused, 1 wt only, 2 bsj only."""

    def __init__(self, tx, bsj, stage=0, lce_mode=0, keep_mate_detail=False, dump_intervals=False):
        self.tx, self.bsj, self.stage, self.detail = tx, bsj, stage, keep_mate_detail
        o = _Opts(lce_mode, 0, stage, int(keep_mate_detail), int(dump_intervals))
        h = ctypes.c_void_p()
        _check(lib().cg_session_create(tx._h if tx else None, bsj._h if bsj else None, ctypes.byref(o), ctypes.byref(h)))
        self._h = h
        self.n_bsj = bsj.n_tx if (bsj is not None and stage != 1) else 0
        self._keep = None

    def close(self):
        if self._h:
            lib().cg_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, r1, r2):
        """Host buffers -> device, kernels queued; returns immediately.  Buffers stay referenced until fetch()."""
        b1, o1 = flatten_reads(r1)
        b2, o2 = flatten_reads(r2)
        if o1.size != o2.size:
            raise CircallError(-1, "the two mate files hold different numbers of reads")
        self._keep = (b1, o1, b2, o2)
        _check(lib().cg_session_submit(self._h, b1.ctypes.data, o1.ctypes.data, b2.ctypes.data, o2.ctypes.data, o1.size - 1))

    def submit_raw(self, p1, po1, p2, po2, n_pairs):
        """Host pointers (e.g. into pinned staging buffers) without any copy on the Python side."""
        _check(lib().cg_session_submit(self._h, p1, po1, p2, po2, n_pairs))

    def submit_fixed_raw(self, p1, p2, n_pairs, read_len):
        """host pointers to n_pairs * read_len characters per mate file (cg_session_submit_fixed)"""
        _check(lib().cg_session_submit_fixed(self._h, p1, p2, n_pairs, read_len))

    def submit_device(self, d_r1, d_off1, d_r2, d_off2, n_pairs, max_read_len):
        _check(lib().cg_session_submit_device(self._h, d_r1, d_off1, d_r2, d_off2, n_pairs, max_read_len))

    def fetch(self, want_lists=True, materialize=True):
        """Wait for the batch.  want_lists: also bring the per-record lists to host memory (inside the
        library); materialize=False skips copying them into numpy arrays (bench.py's end-to-end leg)."""
        r = _BatchResult()
        try:
            _check(lib().cg_session_fetch(self._h, int(want_lists), ctypes.byref(r)))
        finally:
            self._keep = None
        return BatchResult(r, want_lists and materialize, self.detail)

    def run(self, r1, r2, want_lists=True):
        self.submit(r1, r2)
        return self.fetch(want_lists)

    def counters(self, per_bsj=True):
        c = _Counters()
        pb = np.zeros(self.n_bsj, np.uint64) if per_bsj else None
        _check(lib().cg_session_counters(self._h, ctypes.byref(c), pb.ctypes.data if per_bsj and self.n_bsj else None,
                                         self.n_bsj if per_bsj else 0))
        d = {n: getattr(c, n) for n, _ in _Counters._fields_}
        return (d, pb) if per_bsj else d

    def accumulator(self):
        """(device pointer, number of uint64) of [per-BSJ counts | 8 scalar counters] for the all-reduce."""
        p = ctypes.c_void_p()
        n = ctypes.c_uint64()
        _check(lib().cg_session_accumulator(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    @property
    def launches(self):
        return lib().cg_session_launches(self._h)


def allreduce_sessions(sessions):
    """cg_sessions_allreduce: after it every session's accumulator (and counters) holds the total over all of them."""
    arr = (ctypes.c_void_p * len(sessions))(*[s._h for s in sessions])
    _check(lib().cg_sessions_allreduce(arr, len(sessions)))


class DeviceBuffer:
    """Raw device allocation (resident-input leg of bench.py)."""

    def __init__(self, nbytes, device=0):
        self.device, self.nbytes = device, nbytes
        p = ctypes.c_void_p()
        _check(lib().cg_device_alloc(device, nbytes, ctypes.byref(p)))
        self.ptr = p.value

    def upload(self, arr, offset=0):
        a = np.ascontiguousarray(arr)
        _check(lib().cg_device_upload(self.device, self.ptr + offset, a.ctypes.data, a.nbytes))

    def free(self):
        if self.ptr:
            lib().cg_device_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host staging buffer exposed as a numpy uint8 array."""

    def __init__(self, nbytes):
        p = ctypes.c_void_p()
        _check(lib().cg_host_alloc(nbytes, ctypes.byref(p)))
        self.ptr, self.nbytes = p.value, nbytes
        self.array = np.frombuffer((ctypes.c_char * nbytes).from_address(self.ptr), np.uint8)

    def free(self):
        if self.ptr:
            self.array = None
            lib().cg_host_free(self.ptr)
            self.ptr = None


def flush_l2(device=0):
    _check(lib().cg_device_flush_l2(device))


def pack_reads(r1, off1, r2, off2, max_len, general=False, iters=1, device=0):
    """The read-packing kernel on its own (cg_pack_reads): r1 / r2 concatenated ASCII bases (uint8 arrays), off1 / off2
    uint64 offsets (n_pairs + 1).  Returns (packed [2 n_pairs, W2 + WM] uint64, lens uint16, status, ms per launch)."""
    import numpy as np
    r1 = np.ascontiguousarray(r1, dtype=np.uint8); r2 = np.ascontiguousarray(r2, dtype=np.uint8)
    off1 = np.ascontiguousarray(off1, dtype=np.uint64); off2 = np.ascontiguousarray(off2, dtype=np.uint64)
    n = len(off1) - 1
    assert len(off2) == n + 1
    w2 = max(1, (max_len + 31) // 32)
    wt = w2 + (w2 + 1) // 2
    packed = np.zeros((2 * n, wt), dtype=np.uint64)
    lens = np.zeros(2 * n, dtype=np.uint16)
    st = ctypes.c_int(0)
    ms = ctypes.c_double(0.0)
    _check(lib().cg_pack_reads(device, r1.ctypes.data, off1.ctypes.data, r2.ctypes.data, off2.ctypes.data, n, max_len, int(general), iters,
                               packed.ctypes.data, lens.ctypes.data, ctypes.byref(st), ctypes.byref(ms)))
    return packed, lens, st.value, ms.value


def bench_gather(nbytes, per_thread=64, iters=5, device=0, mode=0):
    g = ctypes.c_double()
    _check(lib().cg_bench_gather_mode(device, nbytes, per_thread, iters, mode, ctypes.byref(g)))
    return g.value
