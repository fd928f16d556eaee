"""K1 (read packing) on its own: ms per launch and algorithmic GB/s (200 B in + 2 x WT x 8 + 4 B out per pair at 2x100)
of the all-bases short cut and of the every-character formulation.  python scripts/pack_bench.py [pairs] [read_len]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import circall_b200 as cg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rng = np.random.default_rng(1)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
b1 = acgt[rng.integers(0, 4, size=n * L, dtype=np.uint8)]
b2 = acgt[rng.integers(0, 4, size=n * L, dtype=np.uint8)]
off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
w2 = (L + 31) // 32
wt = w2 + (w2 + 1) // 2
bytes_per_pair = 2 * L + 16 + 2 * wt * 8 + 4
for general in (False, True):
    p, l, s, ms = cg.pack_reads(b1, off, b2, off, L, general=general, iters=10)
    print("k_pack%s: %d pairs 2x%d  %.3f ms per launch  %.0f GB/s algorithmic  (status %d)" % ("_general" if general else "", n, L, ms, n * bytes_per_pair / ms / 1e6, s))
