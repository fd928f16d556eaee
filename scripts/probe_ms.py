"""Resident inputs through several sessions in rotation (kernels of different batches overlap on the device)."""
import argparse, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import circall_b200 as cg
import synth
ap = argparse.ArgumentParser()
ap.add_argument("--genes", type=int, default=60000)
ap.add_argument("--pairs", type=int, default=16_000_000)
ap.add_argument("--L", type=int, default=100)
a = ap.parse_args()
T = synth.Txome(1, a.genes)
gtx = cg.Index.build(T.tx_text); gbsj = cg.Index.build(T.bsj_text)
r1, r2 = T.reads(2, a.pairs, L=a.L, frag_mean=2.5 * a.L, frag_sd=0.25 * a.L, circ_frac=0.01)
n, L = r1.shape
d1 = cg.DeviceBuffer(r1.nbytes + 64); d1.upload(r1)
d2 = cg.DeviceBuffer(r2.nbytes + 64); d2.upload(r2)
for ns, batch in ((1, 4 << 20), (1, 1 << 20), (2, 2 << 20), (2, 1 << 20), (3, 1 << 20), (4, 1 << 20), (4, 1 << 19)):
    off = np.arange(batch + 1, dtype=np.uint64) * L
    doff = cg.DeviceBuffer(off.nbytes); doff.upload(off)
    SS = [cg.Session(gtx, gbsj, stage=0) for _ in range(ns)]
    best = 1e9
    for it in range(3):
        cg.flush_l2()
        t = time.perf_counter()
        pend = []
        for i, b0 in enumerate(range(0, n, batch)):
            nb = min(batch, n - b0)
            s = SS[i % ns]
            if len(pend) == ns:
                pend.pop(0).fetch(want_lists=False)
            s.submit_device(d1.ptr + b0 * L, doff.ptr, d2.ptr + b0 * L, doff.ptr, nb, L)
            pend.append(s)
        for s in pend:
            s.fetch(want_lists=False)
        best = min(best, time.perf_counter() - t)
    print("sessions %d batch %d: %.1f ms -> %.1f M pairs/s" % (ns, batch, best * 1e3, n / best / 1e6), flush=True)
    for s in SS:
        s.close()
