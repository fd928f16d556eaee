"""Committed golden vectors (tests/golden/golden_v1.json, oracle-generated — see make_golden.py for why there is
nothing reference-generated to pin against): the oracle must still reproduce them (CPU tier) and the CUDA path
must match them through the C ABI (GPU tier)."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests.golden import make_golden as mg

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.json")))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_reproduces_golden(name):
    import oracle
    g = GOLD[name]
    T, r1, r2 = mg.world(g["case"])
    assert mg.digest(T.tx_text) == g["tx_text"] and mg.digest(T.bsj_text) == g["bsj_text"] and mg.digest(np.stack([r1, r2])) == g["reads"]
    tx, bsj = oracle.Index.from_text(T.tx_text), oracle.Index.from_text(T.bsj_text)
    assert mg.digest(tx.sa) == g["tx_sa"] and mg.digest(bsj.sa) == g["bsj_sa"]
    assert (tx.n_kmers, bsj.n_kmers) == (g["tx_kmers"], g["bsj_kmers"])
    for mode in (0, 1):
        assert mg.canonical(oracle.Run(tx, bsj, 2, r1, r2, lce_mode=mode)) == g["lce%d" % mode]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_gpu_matches_golden(name):
    import circall_b200 as cg
    g = GOLD[name]
    T, r1, r2 = mg.world(g["case"])
    tx, bsj = cg.Index.build(T.tx_text), cg.Index.build(T.bsj_text)
    assert mg.digest(tx.suffix_array()) == g["tx_sa"] and mg.digest(bsj.suffix_array()) == g["bsj_sa"]
    for mode in (0, 1):
        want = g["lce%d" % mode]
        S = cg.Session(tx, bsj, lce_mode=mode, keep_mate_detail=True)
        B = S.run(r1, r2)
        ctr, per = S.counters()
        assert mg.digest(np.stack([B.exp_pair.astype(np.uint64), B.exp_code.astype(np.uint64)], 1)) == want["exp"]
        assert mg.digest(np.stack([B.mate_flags.astype(np.uint16), B.mate_nhits], 1)) == want["mates"]
        j = np.stack([B.junc[k].astype(np.int64) for k in ("rec", "dir", "query_pos", "len", "hit_pos", "tid")], 1)
        j = j[np.lexsort(j.T[::-1])]
        assert mg.digest(j) == want["junc"] and j[:5].tolist() == want["first_junc"]
        assert mg.digest(B.cand_rec.astype(np.uint64)) == want["cand"]
        eq = B.eq_classes()
        for t in np.nonzero(per)[0]:
            eq[(int(t),)] = int(per[t])
        assert hashlib.sha256(json.dumps(sorted((list(k), v) for k, v in eq.items())).encode()).hexdigest() == want["eq"]
        assert {k: ctr["wt_" + v] for k, v in (("numObserved", "num_observed"), ("numMapped", "num_mapped"), ("numHits", "num_hits"), ("upperBound", "upper_bound"))} == \
               {k: want["wt"][k] for k in ("numObserved", "numMapped", "numHits", "upperBound")}
        assert {k: ctr["bsj_" + v] for k, v in (("numObserved", "num_observed"), ("numMapped", "num_mapped"), ("numHits", "num_hits"), ("upperBound", "upper_bound"))} == \
               {k: want["bsj"][k] for k in ("numObserved", "numMapped", "numHits", "upperBound")}
        S.close()
