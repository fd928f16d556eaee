"""Regenerates tests/golden/golden_v1.json from the oracle.

The reference ships no tests, golden vectors or expected outputs and cannot be built or run here,
so these vectors are ORACLE-generated ("parity unpinned" in the task's terms): they pin the oracle — and through
it the CUDA path — against silent behaviour changes between rounds.  Inputs are regenerated from seeds by
synth/, so only digests and a handful of explicit rows are stored.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = [dict(name="L100", tx_seed=21, genes=150, read_seed=22, pairs=3000, L=100, circ=0.1, n_rate=0.002, junk=0.03),
         dict(name="L75", tx_seed=21, genes=150, read_seed=23, pairs=2000, L=75, circ=0.1, n_rate=0.0, junk=0.0),
         dict(name="L150", tx_seed=21, genes=150, read_seed=24, pairs=2000, L=150, circ=0.1, n_rate=0.004, junk=0.05)]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def canonical(R):
    """Order-free canonical form of one wt->bsj run (multisets sorted)."""
    p, c = R.wt_export()
    fl, nh = R.wt_mates()
    j = R.bsj_junctions()
    j = j[np.lexsort(j.T[::-1])] if len(j) else j
    eq = sorted((list(k), v) for k, v in R.eq_classes(1).items())
    return dict(exp=digest(np.stack([p, c.astype(np.uint64)], 1)), mates=digest(np.stack([fl.astype(np.uint16), nh], 1)),
                junc=digest(j.astype(np.int64)), cand=digest(R.bsj_candidates()), eq=hashlib.sha256(json.dumps(eq).encode()).hexdigest(),
                n_exp=int(len(p)), n_junc=int(len(j)), n_cand=int(len(R.bsj_candidates())), first_junc=j[:5].tolist(),
                wt={k: R.counters(0)[k] for k in ("numObserved", "numMapped", "numHits", "upperBound", "readLen")},
                bsj={k: R.counters(1)[k] for k in ("numObserved", "numMapped", "numHits", "upperBound", "readLen")})


def world(case):
    import synth
    T = synth.Txome(case["tx_seed"], case["genes"], p_antisense=0.05, p_repeat=0.08, p_polya=0.02)
    r1, r2 = T.reads(case["read_seed"], case["pairs"], L=case["L"], frag_mean=2.4 * case["L"], frag_sd=0.3 * case["L"],
                     circ_frac=case["circ"], n_rate=case["n_rate"], junk_frac=case["junk"])
    return T, r1, r2


def main():
    import oracle
    out = {}
    for case in CASES:
        T, r1, r2 = world(case)
        tx, bsj = oracle.Index.from_text(T.tx_text), oracle.Index.from_text(T.bsj_text)
        g = dict(case=case, tx_text=digest(T.tx_text), bsj_text=digest(T.bsj_text), reads=digest(np.stack([r1, r2])),
                 tx_sa=digest(tx.sa), bsj_sa=digest(bsj.sa), tx_kmers=int(tx.n_kmers), bsj_kmers=int(bsj.n_kmers))
        for mode in (0, 1):
            g["lce%d" % mode] = canonical(oracle.Run(tx, bsj, 2, r1, r2, lce_mode=mode))
        out[case["name"]] = g
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
