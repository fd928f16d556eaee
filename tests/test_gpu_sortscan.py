"""Row N2's primitives on their own (circall_b200/csrc/cg_sortscan.cuh, through cg_debug_sort_pairs / cg_debug_scan): the stable radix
sort of (uint64, uint32) pairs, the three-phase scans and the flag selection of the index builder against numpy, at sizes around the
tile boundaries (4096 pairs per sort tile, 4096 entries per scan chunk, 8192 chunk totals per trip of the single-block scan)."""
import numpy as np
import pytest

import circall_b200 as cg

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 31, 32, 33, 511, 4095, 4096, 4097, 8191, 65537, 1_000_003, 4096 * 8192 + 5]


@pytest.mark.parametrize("n", SIZES)
def test_radix_sort_is_a_stable_sort(n):
    rng = np.random.default_rng(n)
    if n > 2_000_000:
        n = 3_000_001                                               # (the largest size is for the scans)
    # few distinct keys in some digits (ties everywhere), all 64 bits in others
    keys = rng.integers(0, 1 << 63, n, dtype=np.uint64) & np.uint64(0xFFFF0000FF00F00F)
    keys[rng.random(n) < 0.3] = keys[0]
    vals = np.arange(n, dtype=np.uint32)
    k, v = cg.debug_sort_pairs(keys, vals, 0, 64)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])
    # a bit range: only those bits decide, everything else keeps its order
    k2, v2 = cg.debug_sort_pairs(keys, vals, 12, 41)
    sub = (keys >> np.uint64(12)) & np.uint64((1 << 29) - 1)
    order = np.argsort(sub, kind="stable")
    assert np.array_equal(k2, keys[order]) and np.array_equal(v2, vals[order])


@pytest.mark.parametrize("n", SIZES)
def test_scans_and_selection(n):
    rng = np.random.default_rng(7 * n + 1)
    d = rng.integers(0, 5, n, dtype=np.uint32)
    want = np.concatenate([[0], np.cumsum(d[:-1], dtype=np.uint64)]).astype(np.uint32)
    assert np.array_equal(cg.debug_scan(d, 0), want)
    h = np.where(rng.random(n) < 0.01, np.arange(n, dtype=np.uint32), 0).astype(np.uint32)      # the suffix-array builder's "group heads"
    assert np.array_equal(cg.debug_scan(h, 1), np.maximum.accumulate(h))
    f = (rng.random(n) < 0.37).astype(np.uint32)
    assert np.array_equal(cg.debug_scan(f, 2), np.flatnonzero(f).astype(np.uint32))
    assert cg.debug_scan(np.zeros(n, np.uint32), 2).size == 0
