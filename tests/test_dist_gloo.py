"""N > 1 path on CPU: world_size-2 gloo.  The hot path shards by read pair with the index replicated and no
data-path collective; the only exchange is one all-reduce(sum) of [per-BSJ counts | 8 scalar counters]
(SURVEY.md §8 e).  Each rank maps its shard with the host emulation of the device code, the accumulators are
all-reduced exactly as bench.py does with NCCL, and the result must equal the oracle on the whole input."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    import synth
    from tests import emu
    T = synth.Txome(31, 80)
    otx, obsj = oracle.Index.from_text(T.tx_text, nthreads=2), oracle.Index.from_text(T.bsj_text, nthreads=2)
    etx, ebsj = emu.EmuIndex(T.tx_text, otx.sa), emu.EmuIndex(T.bsj_text, obsj.sa)
    n, L = 3000, 100
    per = n // world
    r1, r2 = T.reads(32, per, L=L, circ_frac=0.15, first=rank * per, nthreads=2)     # bench.py's sharding: first = rank * n
    off = np.arange(per + 1, dtype=np.uint64) * L
    E = emu.run(etx, ebsj, r1.reshape(-1), off, r2.reshape(-1), off)
    acc = np.zeros(T.n_bsj + 8, np.int64)                                            # same layout as cg_session_accumulator
    p0 = 0
    for ln in E["eq_lens"]:
        if ln == 1:
            acc[int(E["eq_tids"][p0])] += 1
        p0 += int(ln)
    acc[T.n_bsj:] = E["ctr"].astype(np.int64)
    t = torch.from_numpy(acc)
    dist.all_reduce(t)
    if rank == 0:
        a1, a2 = T.reads(32, n, L=L, circ_frac=0.15, nthreads=2)
        R = oracle.Run(otx, obsj, 2, a1, a2, nthreads=2)
        want = np.zeros_like(acc)
        for k, v in R.eq_classes(1).items():
            if len(k) == 1:
                want[k[0]] += v
        ow, ob = R.counters(0), R.counters(1)
        want[T.n_bsj:] = [ow["numObserved"], ow["numMapped"], ow["numHits"], ow["upperBound"], ob["numObserved"], ob["numMapped"], ob["numHits"], ob["upperBound"]]
        q.put(bool(np.array_equal(t.numpy(), want)) and int(want[:T.n_bsj].sum()) > 0)
    dist.destroy_process_group()


def test_sharded_counts_allreduce_world2():
    import oracle
    import synth
    from tests import emu
    oracle.build(); synth.build(); emu.build()          # build once here: two ranks rebuilding a stale library at the same time race on the file
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
