"""GPU parity tests: libcircall_gpu.so (through its C ABI) against the CPU oracle on the same inputs.

Integer / index work: everything is compared bit-exactly (multisets where the reference's own order is
nondeterministic, SURVEY.md §8 c)."""
import numpy as np
import pytest

from tests.helpers import assert_batch_equals_oracle, compare_tier_intervals, make_reads, ragged

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_world(small_world):
    import circall_b200 as cg
    if cg.device_count() == 0:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    T, otx, obsj = small_world
    gtx = cg.Index.build(T.tx_text)
    gbsj = cg.Index.build(T.bsj_text)
    return cg, T, otx, obsj, gtx, gbsj


def test_index_matches_oracle(gpu_world):
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    for o, g in ((otx, gtx), (obsj, gbsj)):
        assert g.text_len == o.text_len and g.n_tx == o.n_tx
        assert np.array_equal(g.suffix_array(), o.sa)          # the suffix array is unique
        assert g.n_kmers == o.n_kmers
        for t in (0, o.n_tx // 2, o.n_tx - 1):
            assert g.tx_len(t) == int(o.lens[t]) and g.tx_offset(t) == int(o.offsets[t])


def _cmp_collect(cg, oidx, gidx, reads, lce_mode):
    res = cg.collect(gidx, reads, strict=True, lce_mode=lce_mode)
    n = reads.shape[0]
    for i in range(n):
        found, f, r, hits = oidx.collect(reads[i].tobytes(), strict=True, lce_mode=lce_mode)
        assert bool(res["found"][i]) == found, i
        if not found:
            continue
        gf = res["fwd_ints"][int(res["fwd_off"][i]):int(res["fwd_off"][i + 1])]
        gr = res["rc_ints"][int(res["rc_off"][i]):int(res["rc_off"][i + 1])]
        assert np.array_equal(gf, f), (i, gf, f)
        assert np.array_equal(gr, r), (i, gr, r)
        a, b = int(res["hits_off"][i]), int(res["hits_off"][i + 1])
        gh = list(zip(res["hit_tid"][a:b].tolist(), res["hit_pos"][a:b].tolist(), [bool(x) for x in res["hit_fwd"][a:b]]))
        assert gh == hits, (i, gh[:4], hits[:4])


@pytest.mark.parametrize("lce_mode", [0, 1])
def test_collect_parity(gpu_world, lce_mode):
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    r1, r2 = make_reads(T, 5, 1500)
    _cmp_collect(cg, otx, gtx, r1, lce_mode)
    _cmp_collect(cg, obsj, gbsj, r2, lce_mode)


def _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, r1, r2, lce_mode=0, dump_intervals=False):
    import oracle
    R = oracle.Run(otx, obsj, 2, r1, r2, lce_mode=lce_mode)
    S = cg.Session(gtx, gbsj, stage=0, lce_mode=lce_mode, keep_mate_detail=True, dump_intervals=dump_intervals)
    B = S.run(r1, r2)
    ctr, per_bsj = S.counters()
    assert_batch_equals_oracle(B, ctr, per_bsj, R)
    S.close()
    return B, R


@pytest.mark.parametrize("lce_mode", [0, 1])
def test_production_tier_intervals(gpu_world, lce_mode):
    """Row A3 through the kernels the pipeline really runs: every collector call of a batch (both mates on txIdx, every exported
    record on bsjIdx) hands out its post-vote interval lists together with the tier that answered it — clean-match kernel, lockstep
    thread tier (lean / two-strand), warp kernel (fast / general) — and each list equals the oracle collector's on the same read.
    (`cg_collect` / test_collect_parity checks the same surface through the stand-alone reference kernel.)"""
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    r1, r2 = make_reads(T, 31, 6000)                       # N-rich: those reads go to the warp kernels
    B, R = _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, r1, r2, lce_mode, dump_intervals=True)
    seen = compare_tier_intervals(B, otx, obsj, r1, r2, lce_mode)
    for tool, tiers in (("wt", (1, 2, 3, 4)), ("bsj", (2, 3, 4))):
        for t in tiers:
            assert seen.get((tool, t), 0) > 0, (tool, t, seen)
    a, b = ragged(r1[:2000], r2[:2000], 5)
    B, R = _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, a, b, lce_mode, dump_intervals=True)
    compare_tier_intervals(B, otx, obsj, a, b, lce_mode)


@pytest.mark.parametrize("L,lce_mode", [(100, 0), (100, 1), (75, 0), (150, 0), (250, 0), (33, 0)])
def test_pipeline_parity(gpu_world, L, lce_mode):
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    r1, r2 = make_reads(T, 100 + L, 20000, L=L, frag_mean=2.4 * L, frag_sd=0.3 * L)
    B, R = _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, r1, r2, lce_mode)
    assert L < 50 or (B.n_export > 0 and B.n_junc > 0 and B.n_cand > 0)


@pytest.mark.parametrize("off", ["CG_NO_THR", "CG_NO_PRE", "CG_NO_THR,CG_NO_PRE"])
def test_pipeline_parity_tier_subsets(gpu_world, off, monkeypatch):
    """Every read through the tiers that remain when the thread-per-read front kernels are switched off one by one (the
    lockstep tier alone in front, the clean-match kernel alone, the warp kernels alone): same outputs; and with all of
    them on (the default) the warp kernels must see only a small part of the reads."""
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    for v in off.split(","):
        monkeypatch.setenv(v, "1")
    r1, r2 = make_reads(T, 511, 20000, n_rate=0.0)
    a, b = ragged(r1, r2, 9)
    _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, a, b)
    B, R = _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, r1, r2, 0)
    if off == "CG_NO_PRE":
        assert 0 < B.n_tier2_wt < 0.3 * 2 * r1.shape[0] and 0 < B.n_tier2_bsj < 0.4 * B.n_export


@pytest.mark.parametrize("env", ["CG_NO_RETRY", "CG_NO_LCP", "CG_NO_SATID", "CG_FM=0", "CG_NO_LCP,CG_NO_SATID,CG_FM=0"])
def test_pipeline_parity_without_the_shortcuts(small_world, env, monkeypatch):
    """The exact shortcuts of round 2, switched off one by one: the second two-strand pass over the lean instantiation's
    other-strand declines (session switch), and indices built without lcp bytes (every suffix of an interval compared with the
    text), without the per-suffix transcript ids, without the seed filter.  Same outputs as the oracle every way."""
    import circall_b200 as cg
    T, otx, obsj = small_world
    for v in env.split(","):
        k, _, val = v.partition("=")
        monkeypatch.setenv(k, val or "1")
    gtx, gbsj = cg.Index.build(T.tx_text), cg.Index.build(T.bsj_text)        # the index switches act at build time
    r1, r2 = make_reads(T, 733, 20000, n_rate=0.0005, circ_frac=0.1)
    B, R = _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, r1, r2, 0)
    assert B.n_export > 0 and B.n_junc > 0 and B.n_cand > 0


@pytest.mark.parametrize("k", [21, 25, 29])
def test_pipeline_other_kmer_lengths(small_world, k):
    """TxIndexer takes -k (odd, <= 31; TxIndexer.cpp:87,208-214): indices, tiers, seed filter (its stage geometry follows k - 16) and the lcp
    extension at other k-mer lengths than the pipeline's 31, interval lists of the production tiers included."""
    import circall_b200 as cg
    import oracle
    T, _, _ = small_world
    otx, obsj = oracle.Index.from_text(T.tx_text, k=k), oracle.Index.from_text(T.bsj_text, k=k)
    gtx, gbsj = cg.Index.build(T.tx_text, k=k), cg.Index.build(T.bsj_text, k=k)
    assert gtx.n_kmers == otx.n_kmers and gbsj.n_kmers == obsj.n_kmers and np.array_equal(gtx.suffix_array(), otx.sa)
    r1, r2 = make_reads(T, 300 + k, 12000, n_rate=0.001, circ_frac=0.1)
    B, R = _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, r1, r2, 0)
    assert B.n_export > 0 and B.n_junc > 0


def test_pipeline_ragged_and_edge_reads(gpu_world):
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    r1, r2 = make_reads(T, 77, 6000, n_rate=0.01)
    # whole-N reads, homopolymers and reads shorter than k
    r1[0, :] = ord("N"); r2[1, :] = ord("A"); r1[2, 40:] = ord("N"); r1[3, :] = ord("n")
    r1[4] = np.frombuffer(bytes(r1[4]).lower(), np.uint8)
    a, b = ragged(r1, r2, 3)
    _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, a, b)


def test_pipeline_repeat_rich_world():
    """Repeat families and antisense genes: SA intervals wider than a warp (and wider than maxInterval), both strands
    surviving the vote, large hit lists (> 200 -> cleared) — the wide-interval branches of the warp kernels and the
    hand-over to the general kernels."""
    import circall_b200 as cg
    import oracle
    import synth
    T = synth.Txome(5, 500, p_antisense=0.15, p_repeat=0.7, p_polya=0.05)
    otx, obsj = oracle.Index.from_text(T.tx_text), oracle.Index.from_text(T.bsj_text)
    gtx, gbsj = cg.Index.build(T.tx_text), cg.Index.build(T.bsj_text)
    assert np.array_equal(gtx.suffix_array(), otx.sa)
    r1, r2 = make_reads(T, 6, 30000, circ_frac=0.15)
    Bd, _ = _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, r1[:3000], r2[:3000], 0, dump_intervals=True)
    seen = compare_tier_intervals(Bd, otx, obsj, r1[:3000], r2[:3000], 0)
    assert seen.get(("wt", 5), 0) > 0 and seen.get(("bsj", 5), 0) > 0, seen        # the general kernels answered some of them
    B, R = _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, r1, r2)
    fl, nh = R.wt_mates()
    assert (B.rec_nhits > 32).sum() > 50 and (B.rec_nhits > 200).sum() > 0      # wide hit lists were exercised
    assert B.n_fallback_wt + B.n_fallback_bsj > 0                               # and so was the general path


def test_pipeline_empty_batch(gpu_world):
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    S = cg.Session(gtx, gbsj)
    B = S.run(np.zeros((0, 100), np.uint8), np.zeros((0, 100), np.uint8))
    assert B.n_export == 0 and B.n_junc == 0
    ctr, per = S.counters()
    assert ctr["wt_num_observed"] == 0 and int(per.sum()) == 0


def test_pipeline_rejects_bad_base_and_long_read(gpu_world):
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    r1, r2 = make_reads(T, 9, 64)
    r1[5, 17] = ord("R")
    S = cg.Session(gtx, gbsj)
    with pytest.raises(cg.CircallError) as e:
        S.run(r1, r2)
    assert e.value.code == -4
    S.close()
    S = cg.Session(gtx, gbsj)
    with pytest.raises(cg.CircallError) as e:
        S.run(np.full((2, 300), ord("A"), np.uint8), np.full((2, 300), ord("C"), np.uint8))
    assert e.value.code == -5


def test_stages_compose(gpu_world):
    """wt-only then bsj-only on the exported records == fused (the Circall.sh file hand-off)."""
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    r1, r2 = make_reads(T, 31, 8000)
    F = cg.Session(gtx, gbsj, stage=0).run(r1, r2)
    W = cg.Session(gtx, None, stage=1).run(r1, r2)
    assert np.array_equal(W.exp_pair, F.exp_pair) and np.array_equal(W.exp_code, F.exp_code)
    e1 = np.where(W.exp_code[:, None] == 1, r1[W.exp_pair], r2[W.exp_pair])
    e2 = np.where(W.exp_code[:, None] == 1, r2[W.exp_pair], r1[W.exp_pair])
    Bs = cg.Session(None, gbsj, stage=2).run(e1, e2)
    key = lambda j: sorted(zip(j["rec"].tolist(), j["dir"].tolist(), j["query_pos"].tolist(), j["len"].tolist(), j["hit_pos"].tolist(), j["tid"].tolist()))
    assert key(Bs.junc) == key(F.junc)
    assert np.array_equal(Bs.cand_rec, F.cand_rec)


def test_sessions_on_different_indices_interleave(gpu_world):
    """Two sessions on different index pairs with batches in flight at the same time on one device: the per-device constant index
    view is handed from one pair to the other in stream order (cg_session.cu::acquire_ix), so both get their own results."""
    import oracle
    import synth
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    T2 = synth.Txome(23, 150)
    g2tx, g2bsj = cg.Index.build(T2.tx_text), cg.Index.build(T2.bsj_text)
    ra = make_reads(T, 61, 12000)
    rb = make_reads(T2, 62, 12000)
    want_a = cg.Session(gtx, gbsj).run(*ra)
    want_b = cg.Session(g2tx, g2bsj).run(*rb)
    A, B = cg.Session(gtx, gbsj), cg.Session(g2tx, g2bsj)
    for _ in range(3):
        A.submit(*ra); B.submit(*rb)
        ga, gb = A.fetch(), B.fetch()
        B.submit(*rb); A.submit(*ra)
        gb2, ga2 = B.fetch(), A.fetch()
        for got, want in ((ga, want_a), (ga2, want_a), (gb, want_b), (gb2, want_b)):
            assert np.array_equal(got.exp_pair, want.exp_pair) and np.array_equal(got.exp_code, want.exp_code)
            assert np.array_equal(np.sort(got.junc, order=["rec", "dir", "query_pos", "len", "hit_pos", "tid"]),
                                  np.sort(want.junc, order=["rec", "dir", "query_pos", "len", "hit_pos", "tid"]))
            assert np.array_equal(got.cand_rec, want.cand_rec) and np.array_equal(got.rec_nhits, want.rec_nhits)
    assert want_a.n_junc > 0 and want_b.n_junc > 0


def test_failed_batch_leaves_no_counts(gpu_world):
    """A batch that fails (here: a character outside ACGTN in its last read) must not leave partial counts in the session's accumulator:
    counters, per-BSJ vector and the all-reduce operand only ever hold whole batches, and the session stays usable."""
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    r1, r2 = make_reads(T, 43, 6000)
    S = cg.Session(gtx, gbsj)
    S.run(r1[:3000], r2[:3000], want_lists=False)
    c0, p0 = S.counters()
    bad = r1[3000:].copy()
    bad[-1, 17] = ord("X")
    with pytest.raises(cg.CircallError):
        S.run(bad, r2[3000:], want_lists=False)
    c1, p1 = S.counters()
    assert c1 == c0 and np.array_equal(p1, p0)
    S.run(r1[3000:], r2[3000:], want_lists=False)
    c2, p2 = S.counters()
    W = cg.Session(gtx, gbsj)
    W.run(r1, r2, want_lists=False)
    cw, pw = W.counters()
    assert c2 == cw and np.array_equal(p2, pw) and cw["bsj_num_observed"] > c0["bsj_num_observed"] > 0


def test_sessions_allreduce(gpu_world, monkeypatch):
    """cg_sessions_allreduce: the accumulators of several sessions (one device here, every visible device when there are more; NCCL
    forced on even for one device) sum to the counters of one session over all the reads, and every session ends up with the total."""
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    monkeypatch.setenv("CG_NCCL_ALWAYS", "1")
    r1, r2 = make_reads(T, 47, 9000)
    W = cg.Session(gtx, gbsj)
    W.run(r1, r2, want_lists=False)
    cw, pw = W.counters()
    nd = cg.device_count()
    idx = [(gtx, gbsj)] + [(cg.Index.build(T.tx_text, device=d), cg.Index.build(T.bsj_text, device=d)) for d in range(1, min(nd, 4))]
    SS = [cg.Session(*idx[i % len(idx)]) for i in range(3 * len(idx))]
    parts = np.array_split(np.arange(r1.shape[0]), len(SS))
    for s, ix in zip(SS, parts):
        s.run(r1[ix], r2[ix], want_lists=False)
    cg.allreduce_sessions(SS)
    for s in SS:
        c, p = s.counters()
        for k in cw:
            if not k.endswith("read_len"):
                assert c[k] == cw[k], (k, c[k], cw[k])
        assert np.array_equal(p, pw)
    assert cw["bsj_num_mapped"] > 0 and pw.sum() > 0


def test_batches_accumulate(gpu_world):
    """Counters are linear in the input: two half batches == one full batch."""
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    r1, r2 = make_reads(T, 41, 10000)
    S1 = cg.Session(gtx, gbsj)
    S1.run(r1, r2, want_lists=False)
    c1, p1 = S1.counters()
    S2 = cg.Session(gtx, gbsj)
    S2.run(r1[:5000], r2[:5000], want_lists=False)
    S2.run(r1[5000:], r2[5000:], want_lists=False)
    c2, p2 = S2.counters()
    assert c1 == c2 and np.array_equal(p1, p2)


def test_merge_pairs_parity(gpu_world):
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    import oracle
    rng = np.random.default_rng(8)
    left, right, ls, rs = [], [], [], []
    for i in range(400):
        def mk():
            n = int(rng.integers(0, 7)) if rng.random() < 0.9 else int(rng.integers(150, 260))
            tids = np.sort(rng.choice(300 if n < 100 else 400, n, replace=False))
            return [(int(t), int(rng.integers(-30, 3000)), bool(rng.integers(0, 2))) for t in tids]
        l, r = mk(), mk()
        if rng.random() < 0.3:
            r = [(t, int(rng.integers(-30, 3000)), bool(rng.integers(0, 2))) for t, _, _ in l]
        left.append(l); right.append(r)
        ls.append(bool(l) or rng.random() < 0.5); rs.append(bool(r) or rng.random() < 0.5)
    joff, joint, tm = cg.merge_pairs(left, right, ls, rs, read_len=100, max_num_hits=200)
    for i in range(400):
        oj, otm = oracle.merge_fuzzy(left[i], right[i], ls[i], rs[i], 100, 200)
        gj = joint[int(joff[i]):int(joff[i + 1])]
        # orphan rows carry no mate fields
        assert np.array_equal(gj[:, [0, 1, 2, 6]], oj[:, [0, 1, 2, 6]]) and otm == bool(tm[i]), i
        paired = oj[:, 6] == 3
        assert np.array_equal(gj[paired], oj[paired])


def _pack_numpy(bases, off, max_len):
    """Numpy restatement of the packed-read record (include/circall_gpu.h: cg_pack_reads): 2-bit codes A 0 C 1 G 2 T 3
    (either case), any other character code 0 + N-mask bit; positions past the read's end code 0 / mask 1."""
    n = len(off) - 1
    w2 = max(1, (max_len + 31) // 32)
    wm = (w2 + 1) // 2
    lut = np.full(256, 4, dtype=np.uint8)
    for ch, c in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
        lut[ch] = c
    isn = np.zeros(256, dtype=bool)
    isn[ord("N")] = isn[ord("n")] = True
    rec = np.zeros((n, w2 + wm), dtype=np.uint64)
    lens = np.zeros(n, dtype=np.uint16)
    status = 0
    for i in range(n):
        s = bases[int(off[i]):int(off[i + 1])]
        if len(s) > max_len:
            status |= 2
            s = s[:max_len]
        lens[i] = len(s)
        c = lut[s]
        if np.any((c == 4) & ~isn[s]):
            status |= 1
        mask = np.ones(64 * wm, dtype=np.uint8)
        mask[:len(s)] = (c == 4)
        code = np.zeros(32 * w2, dtype=np.uint64)
        code[:len(s)] = np.where(c == 4, 0, c)
        sh = (2 * (np.arange(32 * w2) % 32)).astype(np.uint64)
        rec[i, :w2] = np.bitwise_or.reduce((code << sh).reshape(w2, 32), axis=1)
        rec[i, w2:] = np.bitwise_or.reduce((mask.astype(np.uint64) << (np.arange(64 * wm) % 64).astype(np.uint64)).reshape(wm, 64), axis=1)
    return rec, lens, status


@pytest.mark.parametrize("max_len,flavour", [(100, "bases"), (100, "mixed"), (75, "mixed"), (150, "mixed"), (250, "mixed"), (33, "mixed"), (64, "junk")])
def test_pack_kernel_parity(max_len, flavour):
    """K1 on its own: the all-bases short cut, the every-character formulation and the numpy restatement agree bit for bit
    (read lengths 0 .. max_len + 3 at every byte alignment, lower case, N, characters the reference rejects)."""
    import circall_b200 as cg
    rng = np.random.default_rng(100 * max_len + len(flavour))
    n = 3000
    alpha = {"bases": b"ACGT", "mixed": b"ACGTacgtACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTNn", "junk": b"ACGTNacgtnX.*@ \x00\xff\x81\x61BDU"}[flavour]
    alpha = np.frombuffer(alpha, dtype=np.uint8)
    mates = []
    for m in range(2):
        ln = rng.integers(0, max_len + 1, size=n)
        ln[rng.random(n) < 0.5] = max_len                       # half of them full length
        if flavour != "bases":
            ln[rng.random(n) < 0.02] = max_len + 3              # too long: cut, status bit 1
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum(ln)
        b = alpha[rng.integers(0, len(alpha), size=int(off[-1]))]
        if flavour == "mixed":                                   # most reads clean, so that both ways run in every warp
            clean = np.repeat(rng.random(n) < 0.7, ln)
            b = np.where(clean, np.frombuffer(b"ACGT", dtype=np.uint8)[b & 3], b)
        mates.append((np.ascontiguousarray(b), off))
    (b1, o1), (b2, o2) = mates
    fast, lf, sf, _ = cg.pack_reads(b1, o1, b2, o2, max_len, general=False)
    gen, lg, sg, _ = cg.pack_reads(b1, o1, b2, o2, max_len, general=True)
    assert np.array_equal(fast, gen) and np.array_equal(lf, lg) and sf == sg
    r1, l1, s1 = _pack_numpy(b1, o1, max_len)
    r2, l2, s2 = _pack_numpy(b2, o2, max_len)
    assert np.array_equal(fast[0::2], r1) and np.array_equal(fast[1::2], r2)
    assert np.array_equal(lf[0::2], l1) and np.array_equal(lf[1::2], l2)
    assert sf == (s1 | s2)
    if flavour == "bases":
        assert sf == 0


def test_pack_kernel_empty():
    import circall_b200 as cg
    z = np.zeros(1, dtype=np.uint64)
    p, l, s, _ = cg.pack_reads(np.zeros(0, np.uint8), z, np.zeros(0, np.uint8), z, 100)
    assert p.shape == (0, 6) and len(l) == 0 and s == 0


@pytest.mark.parametrize("L", [100, 75])
def test_submit_fixed_equals_submit(gpu_world, L):
    """cg_session_submit_fixed (offsets made on the device) gives what cg_session_submit gives for reads of one length."""
    cg, T, otx, obsj, gtx, gbsj = gpu_world
    r1, r2 = make_reads(T, 77 + L, 6000, L=L)
    b1 = np.ascontiguousarray(r1).reshape(-1)
    b2 = np.ascontiguousarray(r2).reshape(-1)
    assert b1.size == 6000 * L and b2.size == 6000 * L
    Sa = cg.Session(gtx, gbsj)
    A = Sa.run(r1, r2)
    Sb = cg.Session(gtx, gbsj)
    Sb.submit_fixed_raw(b1.ctypes.data, b2.ctypes.data, 6000, L)
    B = Sb.fetch()
    ca, pa = Sa.counters()
    cb, pb = Sb.counters()
    assert ca == cb and np.array_equal(pa, pb)
    assert (A.n_export, A.n_junc, A.n_cand, A.n_eq) == (B.n_export, B.n_junc, B.n_cand, B.n_eq)
    assert sorted(zip(A.exp_pair.tolist(), A.exp_code.tolist())) == sorted(zip(B.exp_pair.tolist(), B.exp_code.tolist()))
    assert sorted(A.junc.tolist()) == sorted(B.junc.tolist())
    assert A.eq_classes() == B.eq_classes()
    # an empty batch and the argument checks
    Sb.submit_fixed_raw(None, None, 0, L)
    assert Sb.fetch().n_pairs == 0
    with pytest.raises(cg.CircallError):
        Sb.submit_fixed_raw(b1.ctypes.data, b2.ctypes.data, 10, 0)
    with pytest.raises(cg.CircallError):
        Sb.submit_fixed_raw(b1.ctypes.data, b2.ctypes.data, 10, 100000)
