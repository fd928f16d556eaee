"""CPU tests of the oracle itself: the un-shipped [U] pieces against brute force on small texts, and
known-answer cases for the quirks of SACollectorCircall.hpp listed in SURVEY.md §8 A3."""
import numpy as np
import pytest

import oracle


def rand_text(rng, n_tx, lo, hi, repeat=True):
    txs = []
    for _ in range(n_tx):
        L = int(rng.integers(lo, hi))
        txs.append("".join("ACGT"[x] for x in rng.integers(0, 4, L)))
    if repeat and n_tx > 2:      # shared segments so that k-mers multi-map
        txs[1] = txs[0][: len(txs[0]) // 2] + txs[1]
        txs[2] = txs[2] + txs[0][len(txs[0]) // 3:]
    return "$".join(txs) + "$", txs


@pytest.fixture(scope="module")
def tiny():
    rng = np.random.default_rng(5)
    text, txs = rand_text(rng, 12, 60, 300)
    ix = oracle.Index.from_text(np.frombuffer(text.encode(), np.uint8), k=15)
    return rng, text, txs, ix


def test_suffix_array_is_sorted(tiny):
    rng, text, txs, ix = tiny
    sa = ix.sa
    suf = [text[p:] for p in sa]
    assert suf == sorted(suf)                       # raw byte order: '$' < 'A' < 'C' < 'G' < 'T'
    assert sorted(sa.tolist()) == list(range(len(text)))


def test_rank_and_offsets(tiny):
    rng, text, txs, ix = tiny
    for p in rng.integers(0, len(text), 300):
        assert ix.tid_at(int(p)) == text[:int(p)].count("$")
    starts = np.cumsum([0] + [len(t) + 1 for t in txs[:-1]])
    assert np.array_equal(ix.offsets, starts) and np.array_equal(ix.lens, [len(t) for t in txs])


def test_khash_is_the_kmer_interval_map(tiny):
    rng, text, txs, ix = tiny
    k = 15
    sa = ix.sa
    kmers = {}
    for i, p in enumerate(sa):
        w = text[p:p + k]
        if len(w) == k and "$" not in w:
            kmers.setdefault(w, []).append(i)
    assert ix.n_kmers == len(kmers)
    for w, idx in list(kmers.items())[::7]:
        assert ix.khash_find(w) == (idx[0], idx[-1] + 1) and idx == list(range(idx[0], idx[-1] + 1))
    assert ix.khash_find("A" * k) is None or "A" * k in kmers


def test_extend_matches_bruteforce(tiny):
    rng, text, txs, ix = tiny
    k = 15
    for _ in range(400):
        t = txs[int(rng.integers(0, len(txs)))]
        a = int(rng.integers(0, len(t) - k - 1))
        L = int(rng.integers(k + 1, min(90, len(t) - a)))
        read = list(t[a:a + L])
        if rng.random() < 0.6:
            e = int(rng.integers(k, L))
            read[e] = "ACGT"[("ACGT".index(read[e]) + 1 + int(rng.integers(0, 3))) % 4]
        if rng.random() < 0.1:
            read[int(rng.integers(k, L))] = "N"
        read = "".join(read)
        be = ix.khash_find(read[:k])
        assert be is not None
        lb = max(0, be[0] - 1)
        assert ix.extend(lb, be[1], k, read, 0) == ix.extend(lb, be[1], k, read, 0, brute=True)
        # reverse-complement query against the same interval semantics
        rc = read[::-1].translate(str.maketrans("ACGTN", "TGCAN"))
        be = ix.khash_find(read[:k])
        got = ix.extend(lb, be[1], k, rc, 0, rc=True)
        assert got == ix.extend(lb, be[1], k, rc, 0, rc=True, brute=True)


def test_lce_modes(tiny):
    rng, text, txs, ix = tiny
    sa = ix.sa
    n = len(text)
    for _ in range(300):
        p1, p2 = (int(x) for x in rng.integers(0, n, 2))
        start = int(rng.integers(0, 20)); stop = int(rng.integers(0, 80))
        for mode in (0, 1):
            o1 = sa[p1] + (start if mode == 0 else 0); o2 = sa[p2] + (start if mode == 0 else 0)
            ln = start
            while max(o1, o2) + ln < n and text[o1 + ln] == text[o2 + ln]:
                if text[o1 + ln] == "$" or ln >= stop:
                    break
                ln += 1
            assert ix.lce(p1, p2, start, stop, mode) == ln


def test_quirk_perfect_read_gives_three_intervals():
    rng = np.random.default_rng(9)
    text, txs = rand_text(rng, 6, 300, 400, repeat=False)
    ix = oracle.Index.from_text(np.frombuffer(text.encode(), np.uint8), k=31)
    read = txs[2][40:140]
    found, f, r, hits = ix.collect(read)
    assert found and len(r) == 0
    assert [tuple(x[2:]) for x in f] == [(100, 0), (31, 69), (31, 69)]      # (len, queryPos): SURVEY A3
    assert hits == [(2, 40, True)]
    # reverse-complement read: queryPos measured on the reversed read
    rc = read[::-1].translate(str.maketrans("ACGT", "TGCA"))
    found, f, r, hits = ix.collect(rc)
    assert found and len(f) == 0 and [tuple(x[2:]) for x in r][0] == (100, 0)
    assert hits == [(2, 40, False)]


def test_quirk_last_kmer_is_never_a_first_seed():
    rng = np.random.default_rng(10)
    text, txs = rand_text(rng, 4, 300, 400, repeat=False)
    ix = oracle.Index.from_text(np.frombuffer(text.encode(), np.uint8), k=31)
    junk = "".join("ACGT"[x] for x in rng.integers(0, 4, 69))
    read = junk + txs[1][100:131]                   # only the last 31-mer matches: Phase A loop is `re < end`
    found, f, r, hits = ix.collect(read)
    assert not found and not hits
    read2 = junk[:68] + txs[1][100:132]             # one base earlier: seeds
    assert ix.collect(read2)[0]


def test_quirk_phase_a_n_guard_is_inclusive():
    rng = np.random.default_rng(11)
    text, txs = rand_text(rng, 4, 300, 400, repeat=False)
    ix = oracle.Index.from_text(np.frombuffer(text.encode(), np.uint8), k=31)
    t = txs[0]
    read = list(t[50:150])
    read[31] = "N"                                   # invalidPos == pos + k  ->  Phase A skips past it (:118)
    found, f, r, hits = ix.collect("".join(read))
    assert found and all(x[3] >= 32 for x in f)
    read = list(t[50:150]); read[32] = "N"           # invalidPos == pos + k + 1 -> first k-mer is used
    found, f, r, hits = ix.collect("".join(read))
    assert found and f[0][3] == 0 and f[0][2] == 32


def test_homopolymer_kmers_are_skipped():
    text = "ACGT" * 20 + "A" * 60 + "GATTACA" * 12 + "$"
    ix = oracle.Index.from_text(np.frombuffer(text.encode(), np.uint8), k=31)
    assert ix.khash_find("A" * 31) is not None
    found, f, r, hits = ix.collect("A" * 80)
    assert not found                                  # every window is a homopolymer: never probed (:127)


def test_std_sort_replica_matches_libstdcxx():
    from tests import emu
    rng = np.random.default_rng(3)
    for n in list(range(0, 40)) + [64, 100, 257, 500, 900]:
        for _ in range(6):
            keys = rng.integers(0, max(2, n // 3 + 1), n).astype(np.uint32)
            a = (keys << 8) | rng.integers(0, 16, n).astype(np.uint32)
            assert np.array_equal(emu.stdsort(a), oracle.stdsort(a)), n


def test_merge_fuzzy_semantics():
    L = [(3, 10, True), (7, -4, False), (9, 500, True)]
    R = [(7, 120, True), (9, 380, False), (11, 5, True)]
    j, tm = oracle.merge_fuzzy(L, R, True, True, 100, 200)
    assert not tm and j[:, 0].tolist() == [7, 9] and j[:, 6].tolist() == [3, 3]
    assert j[0].tolist() == [7, 0, 0, 120, 1, 220, 3]            # pos clamped at 0; fragLen = 120 + 100 - 0
    assert j[1].tolist() == [9, 500, 1, 380, 0, 220, 3]
    j, tm = oracle.merge_fuzzy(L, [(1, 1, True)], True, True)    # nothing in common: all left then all right
    assert j[:, 0].tolist() == [3, 7, 9, 1] and j[:, 6].tolist() == [1, 1, 1, 2]
    j, tm = oracle.merge_fuzzy([], R, False, True)               # left mate never seeded: right orphans
    assert j[:, 0].tolist() == [7, 9, 11]
    j, tm = oracle.merge_fuzzy([], R, True, True)                # left seeded but no hits: dropped
    assert j.shape[0] == 0
