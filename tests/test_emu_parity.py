"""CPU tier: the device code (cg_core.cuh / cg_pipeline.cuh compiled for the host by tests/emu) against the
oracle.  This is how the CUDA logic is checked where no GPU exists; the `-m gpu` tests repeat it through
libcircall_gpu.so on the B200."""
import numpy as np
import pytest

import oracle
from tests import emu
from tests.helpers import make_reads, ragged


@pytest.fixture(scope="module")
def world(small_world):
    T, otx, obsj = small_world
    etx = emu.EmuIndex(T.tx_text, otx.sa)
    ebsj = emu.EmuIndex(T.bsj_text, obsj.sa)
    assert etx.n_kmers == otx.n_kmers and ebsj.n_kmers == obsj.n_kmers
    return T, otx, obsj, etx, ebsj


@pytest.mark.parametrize("lce_mode", [0, 1])
def test_collect_intervals_and_hits(world, lce_mode):
    T, otx, obsj, etx, ebsj = world
    r1, r2 = make_reads(T, 5, 1200, n_rate=0.004)
    for oi, ei, rd in ((otx, etx, r1), (obsj, ebsj, r2)):
        n, L = rd.shape
        res = ei.collect(rd.reshape(-1), np.arange(n + 1, dtype=np.uint64) * L, strict=True, lce_mode=lce_mode)
        for i in range(n):
            found, f, r, hits = oi.collect(rd[i].tobytes(), strict=True, lce_mode=lce_mode)
            assert bool(res["found"][i]) == found
            if not found:
                continue
            assert np.array_equal(res["fwd_ints"][int(res["fwd_off"][i]):int(res["fwd_off"][i + 1])], f)
            assert np.array_equal(res["rc_ints"][int(res["rc_off"][i]):int(res["rc_off"][i + 1])], r)
            a, b = int(res["hits_off"][i]), int(res["hits_off"][i + 1])
            assert list(zip(res["hit_tid"][a:b].tolist(), res["hit_pos"][a:b].tolist(), [bool(x) for x in res["hit_fwd"][a:b]])) == hits


def _pipeline(world, r1, r2, lce_mode=0):
    T, otx, obsj, etx, ebsj = world
    (b1, o1), (b2, o2) = oracle._flat(r1), oracle._flat(r2)
    R = oracle.Run(otx, obsj, 2, r1, r2, lce_mode=lce_mode)
    E = emu.run(etx, ebsj, b1, o1, b2, o2, lce_mode)
    p, c = R.wt_export()
    assert np.array_equal(E["exp_pair"], p) and np.array_equal(E["exp_code"], c)
    fl, nh = R.wt_mates()
    assert np.array_equal(E["mate_flags"], fl) and np.array_equal(E["mate_nhits"], nh)
    assert sorted(map(tuple, R.bsj_junctions().tolist())) == sorted(map(tuple, E["junc"].tolist()))
    assert np.array_equal(E["cand"], R.bsj_candidates())
    onh, osp = R.bsj_records()
    assert np.array_equal(E["rec_nhits"], onh) and np.array_equal(E["rec_split"], osp)
    ow, ob = R.counters(0), R.counters(1)
    assert E["ctr"].tolist() == [ow["numObserved"], ow["numMapped"], ow["numHits"], ow["upperBound"],
                                 ob["numObserved"], ob["numMapped"], ob["numHits"], ob["upperBound"]]
    eq = {}
    p0 = 0
    for ln in E["eq_lens"]:
        key = tuple(int(x) for x in E["eq_tids"][p0:p0 + ln]); p0 += int(ln)
        eq[key] = eq.get(key, 0) + 1
    assert eq == R.eq_classes(1)
    R.emu = E
    return R


@pytest.mark.parametrize("L", [100, 75, 150])
def test_pipeline_lockstep_tier_coverage(world, L, monkeypatch):
    """The lockstep thread tier (cg_thr.cuh) in front of the warp tiers: parity, and it must take most of what the
    clean-match kernel declines (with and without that kernel in front of it)."""
    T = world[0]
    r1, r2 = make_reads(T, 900 + L, 4000, L=L, frag_mean=2.4 * L, frag_sd=0.3 * L)
    R = _pipeline(world, r1, r2, 0)
    ts = R.emu["tiers"]
    pre, thr_wt, thr_bsj = int(ts[0]), int(ts[1]), int(ts[2])
    n_mates, n_rec = 2 * r1.shape[0], len(R.emu["exp_pair"])
    assert thr_wt > 0.4 * (n_mates - pre), (pre, thr_wt, n_mates)       # the test reads are N-rich (18 % of them hold one)
    assert thr_bsj > 0.4 * n_rec, (thr_bsj, n_rec)
    monkeypatch.setenv("CG_EMU_NO_PRE", "1")
    R = _pipeline(world, r1, r2, 1)
    assert int(R.emu["tiers"][1]) > 0.6 * n_mates


@pytest.mark.parametrize("L,lce_mode", [(100, 0), (100, 1), (75, 0), (150, 0), (250, 0), (33, 0)])
def test_pipeline(world, L, lce_mode):
    T = world[0]
    r1, r2 = make_reads(T, 100 + L, 5000, L=L, frag_mean=2.4 * L, frag_sd=0.3 * L)
    R = _pipeline(world, r1, r2, lce_mode)
    assert L < 50 or len(R.bsj_candidates()) > 0


def test_pipeline_ragged_n_rich(world):
    T = world[0]
    r1, r2 = make_reads(T, 77, 3000, n_rate=0.02)
    r1[0, :] = ord("N"); r2[1, :] = ord("A"); r1[2, 40:] = ord("N"); r1[3, :] = ord("n")
    r1[4] = np.frombuffer(bytes(r1[4]).lower(), np.uint8)
    a, b = ragged(r1, r2, 3)
    _pipeline(world, a, b)


def test_pipeline_shipped_style_reads(world):
    """Reads shaped like the reference's shipped fixture (V1/Data/test_read*.fasta.gz: 2x100, ~12 % of reads hold
    an N, a few reads of 99/76/80/82 bases): N-containing windows are "oracle-defined" (SURVEY B5)."""
    T = world[0]
    r1, r2 = make_reads(T, 55, 3000, n_rate=0.0013, circ_frac=0.5)
    rng = np.random.default_rng(1)
    out = []
    for r in (r1, r2):
        n, L = r.shape
        lens = np.full(n, L)
        lens[rng.random(n) < 0.045] = 99
        lens[rng.random(n) < 0.001] = rng.choice([76, 80, 82])
        off = np.zeros(n + 1, np.uint64); off[1:] = np.cumsum(lens)
        out.append((np.concatenate([r[i, :lens[i]] for i in range(n)]), off))
    _pipeline(world, out[0], out[1])


def test_seed_filter_exact_and_cheaper(small_world, monkeypatch):
    """The lockstep tier's walks with the index's seed filter (cg_thr.cuh::q_walk; which fm-mers occur in the text) against the
    same walks without it: the outputs are those of the oracle either way — the filter may only remove windows that are not in
    the table — and with the filter the walks look up far fewer table slots (each one a random DRAM access on the device).
    fm = 12 sets a sixth of its bits on this text instead of a few in ten thousand at 16: the "present" side of every branch runs too."""
    T, otx, obsj = small_world
    r1, r2 = T.reads(77, 6000, L=100, circ_frac=0.05)           # no N: every read reaches the thread tiers
    cost = {}
    for fm in (0, 12, 16):
        monkeypatch.setenv("CG_FM", str(fm))
        etx, ebsj = emu.EmuIndex(T.tx_text, otx.sa), emu.EmuIndex(T.bsj_text, obsj.sa)
        assert etx.fm == fm and ebsj.fm == fm
        emu.walk_stats()
        _pipeline((T, otx, obsj, etx, ebsj), r1, r2, 0)
        cost[fm] = emu.walk_stats()
    assert cost[0][0] == 0 and cost[0][2] == 0
    for fm in (12, 16):
        filt, table, dead = cost[fm]
        assert dead > 0 and filt > 0
        assert filt + table < 0.5 * cost[0][1], cost              # random accesses of the walks: at least halved
    assert cost[16][1] < 0.35 * cost[0][1], cost


def test_lcp_extension_exact(small_world, monkeypatch):
    """extendSearchNaive over a k-mer interval with the suffix array's lcp bytes (cg_thr.cuh::q_step, cg_pre.cuh: the match length of a
    suffix from that of the one before it and the lcp of the two; the text is read for the first suffix and where the two are equal)
    against the same tiers comparing every suffix with the text: the outputs are the oracle's either way, and the lcp path compares far
    fewer 32-base words."""
    T, otx, obsj = small_world
    r1, r2 = T.reads(78, 5000, L=100, circ_frac=0.05)
    L = emu._lib()
    import ctypes
    L.emu_cost_log.restype = ctypes.c_uint64
    L.emu_cost_log.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    words = {}
    for no_lcp in (False, True):
        if no_lcp:
            monkeypatch.setenv("CG_NO_LCP", "1")
        etx, ebsj = emu.EmuIndex(T.tx_text, otx.sa), emu.EmuIndex(T.bsj_text, obsj.sa)
        for w in (0, 1):
            L.emu_cost_log(w, None, 1)
        _pipeline((T, otx, obsj, etx, ebsj), r1, r2, 0)
        tot = 0
        for w in (0, 1):
            n = L.emu_cost_log(w, None, 0)
            a = np.zeros(n, np.uint32)
            L.emu_cost_log(w, a.ctypes.data, 1)
            tot += int(a.reshape(-1, 10)[:, 2 + 4].sum())        # q_match word steps
        words[no_lcp] = tot
    assert 0 < words[False] < 0.6 * words[True], words


@pytest.mark.parametrize("k", [21, 25, 29])
def test_pipeline_other_kmer_lengths(small_world, k):
    """TxIndexer takes -k (odd, <= 31; TxIndexer.cpp:87,208-214).  The tiers, the seed filter (stage geometry follows k - 16) and the lcp
    extension at other k-mer lengths than the pipeline's 31."""
    T, _, _ = small_world
    otx, obsj = oracle.Index.from_text(T.tx_text, k=k), oracle.Index.from_text(T.bsj_text, k=k)
    etx, ebsj = emu.EmuIndex(T.tx_text, otx.sa, k), emu.EmuIndex(T.bsj_text, obsj.sa, k)
    assert etx.n_kmers == otx.n_kmers and ebsj.n_kmers == obsj.n_kmers
    r1, r2 = make_reads(T, 300 + k, 2500, n_rate=0.001, circ_frac=0.1)
    _pipeline((T, otx, obsj, etx, ebsj), r1, r2, 0)
