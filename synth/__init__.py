"""Synthetic transcriptome / BSJ reference / read-pair generator (tests + bench.py only).

Thin ctypes wrapper over synth/synth.cpp (built into synth/libsynth.so by __graft_entry__.build()).
Neither the product (circall_b200/) nor the oracle depends on it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libsynth.so")
    src = os.path.join(_HERE, "synth.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-fPIC", "-shared", "-pthread",
                               "-o", so, src])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libsynth.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.syn_txome_create.restype = ctypes.c_void_p
        L.syn_txome_create.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_double, ctypes.c_double, ctypes.c_uint32]
        L.syn_txome_free.argtypes = [ctypes.c_void_p]
        for f in (L.syn_tx_text, L.syn_bsj_text):
            f.restype = ctypes.c_void_p
            f.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)]
        L.syn_num_genes.restype = ctypes.c_uint32
        L.syn_num_genes.argtypes = [ctypes.c_void_p]
        for f in (L.syn_tx_name, L.syn_bsj_name):
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_int]
        L.syn_gene_exons.restype = ctypes.c_int
        L.syn_gene_exons.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.syn_reads.restype = None
        L.syn_reads.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                                ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _LIB = L
    return _LIB


class Txome:
    """A synthetic gene set with its transcript text and BSJ-reference text ('$'-joined, uint8)."""

    def __init__(self, seed, n_genes, mean_iso=3.3, p_antisense=0.02, p_repeat=0.04, p_polya=0.005, bsj_flank=149):
        L = _lib()
        self._h = L.syn_txome_create(seed, n_genes, mean_iso, p_antisense, p_repeat, p_polya, bsj_flank)
        n = ctypes.c_uint64()
        m = ctypes.c_uint32()
        p = L.syn_tx_text(self._h, ctypes.byref(n), ctypes.byref(m))
        self.tx_text = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(n.value,))
        self.n_tx = m.value
        p = L.syn_bsj_text(self._h, ctypes.byref(n), ctypes.byref(m))
        self.bsj_text = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(n.value,))
        self.n_bsj = m.value
        self.n_genes = n_genes

    def close(self):
        if self._h:
            _lib().syn_txome_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tx_name(self, t):
        b = ctypes.create_string_buffer(64)
        _lib().syn_tx_name(self._h, t, b, 64)
        return b.value.decode()

    def bsj_name(self, i):
        b = ctypes.create_string_buffer(128)
        _lib().syn_bsj_name(self._h, i, b, 128)
        return b.value.decode()

    def gene_exons(self, g):
        """[(start, end)] of gene g's exons in the coordinates of the BSJ reference names (1-based, inclusive)"""
        st = np.zeros(32, np.uint64)
        en = np.zeros(32, np.uint64)
        n = _lib().syn_gene_exons(self._h, g, st.ctypes.data, en.ctypes.data, 32)
        return [(int(st[i]), int(en[i])) for i in range(n)]

    def reads(self, seed, n_pairs, L=100, frag_mean=250.0, frag_sd=25.0, err=0.005, circ_frac=0.01, n_rate=0.0,
              junk_frac=0.0, first=0, truth=False, nthreads=None, out=None):
        """Return (r1, r2[, truth]) as (n_pairs, L) uint8 arrays of ASCII bases."""
        if nthreads is None:
            nthreads = os.cpu_count() or 1
        if out is None:
            r1 = np.empty((n_pairs, L), dtype=np.uint8)
            r2 = np.empty((n_pairs, L), dtype=np.uint8)
        else:
            r1, r2 = out
        tr = np.empty((n_pairs, 3), dtype=np.int32) if truth else None
        _lib().syn_reads(self._h, seed, first, n_pairs, L, frag_mean, frag_sd, err, circ_frac, n_rate, junk_frac,
                         r1.ctypes.data, r2.ctypes.data, tr.ctypes.data if truth else None, nthreads)
        return (r1, r2, tr) if truth else (r1, r2)


# BASELINE.json configs (SURVEY.md §8 d).  n_genes is chosen so the transcript count lands near
# the named figure with the default mean_iso.
CONFIGS = {
    1: dict(name="cfg1: ~5k-tx synthetic transcriptome + BSJ ref, 1M 2x100 pairs", tx_seed=1, read_seed=2,
            n_genes=1500, n_pairs=1_000_000, L=100, frag_mean=250.0, frag_sd=25.0, circ_frac=0.01),
    2: dict(name="cfg2: hg19-scale synthetic transcriptome (~200k tx) + BSJ index, 20M 2x100 pairs", tx_seed=3,
            read_seed=4, n_genes=60000, n_pairs=20_000_000, L=100, frag_mean=250.0, frag_sd=25.0, circ_frac=0.01),
    3: dict(name="cfg3: hg38-scale synthetic transcriptome (~250k tx) + BSJ index, 100M 2x150 pairs over 8 GPUs",
            tx_seed=5, read_seed=6, n_genes=75000, n_pairs=100_000_000, L=150, frag_mean=300.0, frag_sd=30.0,
            circ_frac=0.01),
    4: dict(name="cfg4: cfg2 index, 50M 2x100 pairs, 10% back-spliced fragments", tx_seed=3, read_seed=7,
            n_genes=60000, n_pairs=50_000_000, L=100, frag_mean=250.0, frag_sd=25.0, circ_frac=0.10),
    5: dict(name="cfg5: cfg2 index, 200M 2x75 pairs, variable fragment length (short anchors)", tx_seed=3,
            read_seed=8, n_genes=60000, n_pairs=200_000_000, L=75, frag_mean=180.0, frag_sd=60.0, circ_frac=0.01),
}
