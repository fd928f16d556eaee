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
