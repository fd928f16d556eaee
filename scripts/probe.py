"""Quick GPU probe: build a synthetic index, run the fused session on resident inputs, print timings."""
import argparse
import json
import sys
import time
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import circall_b200 as cg
import synth

ap = argparse.ArgumentParser()
ap.add_argument("--genes", type=int, default=1500)
ap.add_argument("--pairs", type=int, default=1_000_000)
ap.add_argument("--batch", type=int, default=1 << 20)
ap.add_argument("--L", type=int, default=100)
ap.add_argument("--circ", type=float, default=0.01)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--gather", action="store_true")
a = ap.parse_args()

t = time.time(); T = synth.Txome(1, a.genes); print("txome %.1fs tx=%d (%d b) bsj=%d (%d b)" % (time.time() - t, T.n_tx, T.tx_text.size, T.n_bsj, T.bsj_text.size), flush=True)
t = time.time(); gtx = cg.Index.build(T.tx_text); print("tx index %.2fs kmers=%d bytes=%.2f GB" % (time.time() - t, gtx.n_kmers, gtx.device_bytes / 1e9), flush=True)
t = time.time(); gbsj = cg.Index.build(T.bsj_text); print("bsj index %.2fs kmers=%d bytes=%.2f GB" % (time.time() - t, gbsj.n_kmers, gbsj.device_bytes / 1e9), flush=True)
t = time.time(); r1, r2 = T.reads(2, a.pairs, L=a.L, frag_mean=2.5 * a.L, frag_sd=0.25 * a.L, circ_frac=a.circ); print("reads %.1fs" % (time.time() - t), flush=True)
n, L = r1.shape
d1 = cg.DeviceBuffer(r1.nbytes + 64); d1.upload(r1)
d2 = cg.DeviceBuffer(r2.nbytes + 64); d2.upload(r2)
off = np.arange(a.batch + 1, dtype=np.uint64) * L
doff = cg.DeviceBuffer(off.nbytes); doff.upload(off)
for stage in (1, 0):
    S = cg.Session(gtx, gbsj if stage == 0 else None, stage=stage)
    for it in range(a.iters):
        cg.flush_l2()
        tot = 0.0; nexp = 0; njunc = 0; fb = [0, 0]; t2 = [0, 0]; km = [0.0] * 4
        w = time.time()
        for b0 in range(0, n, a.batch):
            nb = min(a.batch, n - b0)
            S.submit_device(d1.ptr + b0 * L, doff.ptr, d2.ptr + b0 * L, doff.ptr, nb, L)
            B = S.fetch(want_lists=False)
            tot += B.device_ms; nexp += B.n_export; njunc += B.n_junc; fb[0] += B.n_fallback_wt; fb[1] += B.n_fallback_bsj; t2[0] += B.n_tier2_wt; t2[1] += B.n_tier2_bsj; km = [a + b for a, b in zip(km, B.kernel_ms)]
        w = time.time() - w
        print("stage %d iter %d: device %.2f ms wall %.2f ms  -> %.2f M pairs/s (device)  export=%d junc=%d fallback wt/bsj=%d/%d tier2 wt/bsj=%d/%d kernel ms pack/wt/commit/bsj=%s" % (stage, it, tot, w * 1e3, n / tot / 1e3, nexp, njunc, fb[0], fb[1], t2[0], t2[1], ["%.1f" % x for x in km]), flush=True)
    print(S.counters(per_bsj=False), "launches", S.launches)
    S.close()
if a.gather:
    for gb in (0.25, 4, 16):
        print("gather %.2f GB: %.0f GB/s" % (gb, cg.bench_gather(int(gb * (1 << 30)), 64, 3)))
