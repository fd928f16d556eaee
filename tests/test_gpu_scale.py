"""Parity at BASELINE.json's scale: the same full comparison the small-world tests make — export list, per-mate
flags and hit counts, junction-row multiset, candidate list, per-record hits / split flags, all eight counters, equivalence classes
and the per-BSJ count vector — on

  * config 1 whole: 1 500 genes (~5 k transcripts), 1 M 2x100 pairs; the oracle builds its own suffix arrays here (induced sorting,
    oracle/u_sais.cpp) and the GPU builder's arrays must equal them;
  * 400 k-pair samples of configs 2, 4 and 5 on the hg19-scale index (60 000 genes) and of config 3 (2x150) on the hg38-scale index
    (75 000 genes).  At this scale the oracle takes the GPU-built suffix array after verifying it (permutation + suffix order, in
    linear time): the array is unique, so a verified one is the one any sorter would return;

plus row A3 on the production tiers (post-vote interval lists per collector call, tests/helpers.py::compare_tier_intervals) on a
sub-sample that always includes every call the warp-per-read kernels answered.  Integer work: everything is compared bit-exactly.
"""
import os

import numpy as np
import pytest

from tests.helpers import compare_tier_intervals
from tests.test_gpu_pipeline import _cmp_pipeline

pytestmark = pytest.mark.gpu

SAMPLE = 400_000


def _reads(T, cfg, n):
    return T.reads(cfg["read_seed"], n, L=cfg["L"], frag_mean=cfg["frag_mean"], frag_sd=cfg["frag_sd"], circ_frac=cfg["circ_frac"])


def _tier_items(B, every):
    """collector calls to check interval by interval: every `every`-th one, and all that a warp-per-read kernel answered"""
    wt_tier = (B.wt_iv[0] >> 24) & 0xf
    wt = sorted(set(range(0, wt_tier.size, every)) | set(np.nonzero(wt_tier >= 4)[0][:4000].tolist()))
    bs_tier = (B.bsj_iv[0] >> 24) & 0xf
    bs = sorted(set(range(0, bs_tier.size, every)) | set(np.nonzero(bs_tier >= 4)[0][:4000].tolist()))
    return wt, bs


def test_config1_whole():
    import circall_b200 as cg
    import oracle
    import synth
    if cg.device_count() == 0:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    cfg = synth.CONFIGS[1]
    T = synth.Txome(cfg["tx_seed"], cfg["n_genes"])
    otx, obsj = oracle.Index.from_text(T.tx_text), oracle.Index.from_text(T.bsj_text)          # the oracle's own suffix sort
    gtx, gbsj = cg.Index.build(T.tx_text), cg.Index.build(T.bsj_text)
    assert np.array_equal(gtx.suffix_array(), otx.sa) and np.array_equal(gbsj.suffix_array(), obsj.sa)
    assert gtx.n_kmers == otx.n_kmers and gbsj.n_kmers == obsj.n_kmers
    r1, r2 = _reads(T, cfg, cfg["n_pairs"])
    B, R = _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, r1, r2, 0, dump_intervals=True)
    assert B.n_export > 100_000 and B.n_junc > 1000 and B.n_cand > 1000
    wt, bs = _tier_items(B, 97)
    seen = compare_tier_intervals(B, otx, obsj, r1, r2, 0, wt, bs)
    assert all(seen.get(("wt", t), 0) > 0 for t in (1, 2, 3, 4)) and all(seen.get(("bsj", t), 0) > 0 for t in (2, 3, 4)), seen


@pytest.fixture(scope="module", params=["hg19-scale", "hg38-scale"])
def scale_world(request):
    """(cg, T, oracle indices, GPU indices, configs that run on this index); one index pair resident at a time"""
    import circall_b200 as cg
    import oracle
    import synth
    if cg.device_count() == 0:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    cfgs = [2, 4, 5] if request.param == "hg19-scale" else [3]
    c0 = synth.CONFIGS[cfgs[0]]
    T = synth.Txome(c0["tx_seed"], c0["n_genes"])
    gtx, gbsj = cg.Index.build(T.tx_text), cg.Index.build(T.bsj_text)
    cores = os.cpu_count() or 1
    otx = oracle.Index.from_text_with_sa(T.tx_text, gtx.suffix_array(), nthreads=cores)     # verified before use (oracle/u_index.cpp::verify_sa)
    obsj = oracle.Index.from_text_with_sa(T.bsj_text, gbsj.suffix_array(), nthreads=cores)
    assert gtx.n_kmers == otx.n_kmers and gbsj.n_kmers == obsj.n_kmers
    yield cg, T, otx, obsj, gtx, gbsj, cfgs
    gtx.close(); gbsj.close(); otx.close(); obsj.close(); T.close()


def test_baseline_config_samples(scale_world):
    import synth
    cg, T, otx, obsj, gtx, gbsj, cfgs = scale_world
    for c in cfgs:
        cfg = synth.CONFIGS[c]
        r1, r2 = _reads(T, cfg, SAMPLE)
        B, R = _cmp_pipeline(cg, T, otx, obsj, gtx, gbsj, r1, r2, 0, dump_intervals=True)
        assert B.n_export > 0.1 * SAMPLE and B.n_junc > 100 and B.n_cand > 100, (c, B.n_export, B.n_junc, B.n_cand)
        if c == 4:
            assert B.n_cand > 10_000, B.n_cand             # 10 % back-spliced fragments
        wt, bs = _tier_items(B, 199)
        seen = compare_tier_intervals(B, otx, obsj, r1, r2, 0, wt, bs)
        assert all(seen.get(("wt", t), 0) > 0 for t in (1, 2, 3, 4)) and all(seen.get(("bsj", t), 0) > 0 for t in (2, 4)), (c, seen)
