"""Shared helpers of the test modules."""
import numpy as np


def make_reads(T, seed, n, L=100, **kw):
    kw.setdefault("circ_frac", 0.08)
    kw.setdefault("n_rate", 0.002)
    kw.setdefault("junk_frac", 0.03)
    return T.reads(seed, n, L=L, **kw)


def ragged(r1, r2, seed):
    """Trim every read to a random length (ragged input): returns ((bases, off), (bases, off))."""
    rng = np.random.default_rng(seed)
    out = []
    for r in (r1, r2):
        n, L = r.shape
        lens = rng.integers(0, L + 1, n)
        lens[rng.random(n) < 0.5] = L
        off = np.zeros(n + 1, np.uint64)
        off[1:] = np.cumsum(lens)
        b = np.concatenate([r[i, :lens[i]] for i in range(n)]) if n else np.zeros(0, np.uint8)
        out.append((b, off))
    return out[0], out[1]


def compare_tier_intervals(B, oidx_tx, oidx_bsj, r1, r2, lce_mode=0, wt_items=None, bsj_items=None):
    """Row A3 on the PRODUCTION tiers: the post-vote fwdSAInts / rcSAInts (SACollectorCircall.hpp:442-490) of collector calls of a
    batch run with dump_intervals — as the tier that answered the read computed them — against the oracle's collector on the same
    read.  r1 / r2: (n, L) arrays or (bases, offsets).  wt_items: mate indices (2 * pair + mate) to check, default all; bsj_items:
    export records, default all.  Returns {tier: number of collector calls checked}."""
    import numpy as np

    def read_of(reads, i):
        if isinstance(reads, tuple):
            b, o = reads
            return bytes(b[int(o[i]):int(o[i + 1])])
        return reads[i].tobytes()

    seen = {}
    n_mates = B.wt_iv[0].size if B.wt_iv is not None else 0
    for r in (range(n_mates) if wt_items is None else wt_items):
        found, tier, gf, gr = B.intervals_of(B.wt_iv, int(r))
        assert tier != 0, ("no tier recorded mate", r)
        ofound, f, rc, _ = oidx_tx.collect(read_of(r2 if r & 1 else r1, int(r) >> 1), strict=True, lce_mode=lce_mode, cap_hits=1 << 15)
        assert found == ofound, (r, tier, found, ofound)
        if found:
            assert np.array_equal(gf, f) and np.array_equal(gr, rc), (r, tier, gf, f, gr, rc)
        seen[("wt", tier)] = seen.get(("wt", tier), 0) + 1
    if B.bsj_iv is not None:
        for rec in (range(B.n_export) if bsj_items is None else bsj_items):
            found, tier, gf, gr = B.intervals_of(B.bsj_iv, int(rec))
            assert tier != 0, ("no tier recorded record", rec)
            pair, code = int(B.exp_pair[rec]), int(B.exp_code[rec])
            ofound, f, rc, _ = oidx_bsj.collect(read_of(r2 if code == 2 else r1, pair), strict=True, lce_mode=lce_mode, cap_hits=1 << 15)
            assert found == ofound, (rec, tier, found, ofound)
            if found:
                assert np.array_equal(gf, f) and np.array_equal(gr, rc), (rec, tier, gf, f, gr, rc)
            seen[("bsj", tier)] = seen.get(("bsj", tier), 0) + 1
    return seen


def assert_batch_equals_oracle(B, ctr, per_bsj, R):
    """The full comparison of one fused batch with the oracle's run of the same pairs (set equality where the reference's own order
    is free): export list, per-mate flags / hit counts when the session kept them, junction rows, candidate list, per-record hits and
    split flags, the eight scalar counters + read lengths, the equivalence-class map and with it the per-BSJ count vector.
    B: circall_b200.BatchResult fetched with lists; ctr, per_bsj: Session.counters() after that single batch; R: oracle.Run(mode 2)."""
    import numpy as np
    p, c = R.wt_export()
    assert np.array_equal(B.exp_pair, p.astype(np.uint32)) and np.array_equal(B.exp_code, c), "export list"
    if getattr(B, "mate_flags", None) is not None:
        fl, nh = R.wt_mates()
        assert np.array_equal(B.mate_flags, fl) and np.array_equal(B.mate_nhits, nh), "per-mate flags / hit counts"
    oj = R.bsj_junctions()
    gj = np.stack([B.junc[k].astype(np.int64) for k in ("rec", "dir", "query_pos", "len", "hit_pos", "tid")], 1) if B.n_junc else np.zeros((0, 6), np.int64)
    assert oj.shape == gj.shape, "junction row count"
    if oj.shape[0]:
        so, sg = oj[np.lexsort(oj.T[::-1])], gj[np.lexsort(gj.T[::-1])]
        assert np.array_equal(so, sg), "junction rows (as a multiset)"
    assert np.array_equal(B.cand_rec, R.bsj_candidates().astype(np.uint32)), "candidate list"
    onh, osp = R.bsj_records()
    assert np.array_equal(B.rec_nhits, onh) and np.array_equal(B.rec_split, osp), "per-record hits / split"
    ow, ob = R.counters(0), R.counters(1)
    assert (ctr["wt_num_observed"], ctr["wt_num_mapped"], ctr["wt_num_hits"], ctr["wt_upper_bound"]) == \
           (ow["numObserved"], ow["numMapped"], ow["numHits"], ow["upperBound"]), "wt counters"
    assert (ctr["bsj_num_observed"], ctr["bsj_num_mapped"], ctr["bsj_num_hits"], ctr["bsj_upper_bound"]) == \
           (ob["numObserved"], ob["numMapped"], ob["numHits"], ob["upperBound"]), "bsj counters"
    assert ctr["wt_read_len"] == ow["readLen"] and ctr["bsj_read_len"] == ob["readLen"], "read lengths"
    oeq = R.eq_classes(1)
    geq = B.eq_classes()
    for t in np.nonzero(per_bsj)[0]:
        geq[(int(t),)] = int(per_bsj[t])
    assert geq == oeq, "equivalence classes / per-BSJ counts"
