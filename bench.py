#!/usr/bin/env python
"""bench.py — read-pairs/s through the fused Circall_wt + Circall_bsj hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A *step* is one pass of the hot path over the whole synthetic workload (BASELINE.json configs[1]:
hg19-scale synthetic transcriptome + BSJ index, 20 M 2x100 bp pairs per GPU), issued to the library in
batches.  Three legs are measured by the default arm:
  value        inputs resident in HBM, device-timed bracket around K steps
  e2e          the same work through the C ABI with HOST (pinned) buffers (four sessions in rotation): host->device copies of the reads
               and device->host copies of the result lists are inside the timed region
  cpu_baseline the oracle (CPU restatement of the reference algorithm) on a bounded sample, rank 0, N = 1
`--impl reference` times only the CPU restatement (the reference binary cannot be built here).
Nothing here reads /root/reference.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner and, with NCCL_DEBUG set,
# everything else to stdout), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the real stdout.
_REAL_STDOUT = None


def _claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "read-pairs/sec quasi-mapped+BSJ-filtered (fused Circall_wt+Circall_bsj)"
UNIT = "read-pairs/s"


def workload(args):
    import synth
    cfg = dict(synth.CONFIGS[args.config])
    if args.genes:
        cfg["n_genes"] = args.genes
    if args.pairs:
        cfg["n_pairs"] = args.pairs
    return cfg


def config_json(cfg, args):
    """The workload, and nothing that depends on which arm runs it: both arms print the same dictionary (run-dependent figures go
    under `details`)."""
    return {"workload": cfg["name"], "baseline_config_index": args.config - 1, "n_genes": cfg["n_genes"],
            "pairs_per_gpu_per_step": cfg["n_pairs"], "read_len": cfg["L"], "circ_frac": cfg["circ_frac"], "error_rate": 0.005,
            "fragment_mean_sd": [cfg["frag_mean"], cfg["frag_sd"]], "tx_seed": cfg["tx_seed"], "read_seed": cfg["read_seed"],
            "index": "replicated per GPU", "l2": "inputs (>= 4 GB) and index (>= 10 GB) exceed the 126 MB L2; no flush needed"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.p, self.path = gpu, None, "/tmp/cg_clocks_%d_%d.csv" % (os.getpid(), gpu)

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def algorithmic_bytes(c, L, n_junc_rows):
    """bytes = 16 P + w S + T/4 + (w + 8 + 4) H per collector call, + 2 L/4 per mapped read
    (packed read written once, read once), + emitted rows; w = 4 (int32 suffix array).  c = oracle counters."""
    return 16 * c["P"] + 4 * c["S"] + c["T"] / 4 + (4 + 8 + 4) * c["H"] + c["calls"] * (2 * L / 4) + 24 * n_junc_rows


def _oracle_own_indices(oracle, T, cores):
    """The oracle's own indices: suffix arrays by its induced-sorting builder (oracle/u_sais.cpp), the two texts in parallel.  Nothing of
    the product is loaded.  A suffix array is kept under /tmp between runs on the same box (the driver runs this arm once per N); a
    cached array is verified by the oracle (permutation + suffix order) before use, like any array it did not sort itself."""
    import hashlib
    import threading
    out, err = {}, []

    def one(name, text):
        try:
            key = hashlib.sha1(text[:1 << 20].tobytes() + str(text.size).encode()).hexdigest()[:16]
            path = "/tmp/cg_oracle_sa_%s_%s.npy" % (name, key)
            ix = None
            if os.path.exists(path):
                try:
                    ix = oracle.Index.from_text_with_sa(text, np.load(path), nthreads=max(1, cores // 2))
                except Exception:
                    ix = None
            if ix is None:
                ix = oracle.Index.from_text(text, nthreads=max(1, cores // 2))
                try:
                    np.save(path, ix.sa)
                except Exception:
                    pass
            out[name] = ix
        except Exception as e:                                   # surfaced by the caller
            err.append(e)
    th = [threading.Thread(target=one, args=("tx", T.tx_text)), threading.Thread(target=one, args=("bsj", T.bsj_text))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if err:
        raise err[0]
    return out["tx"], out["bsj"]


def reference_arm(args):
    """--impl reference: the CPU restatement of the reference algorithm (oracle/) on all host cores, a bounded sample of the same
    workload per step.  It is NOT the Circall binary (which cannot be built here) and it loads nothing of the
    product: its indices come from the oracle's own suffix sort."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    import synth
    cfg = workload(args)
    cores = os.cpu_count() or 1
    oracle.build(); synth.build()
    T = synth.Txome(cfg["tx_seed"], cfg["n_genes"])
    t0 = time.time()
    otx, obsj = _oracle_own_indices(oracle, T, cores)
    t_index = time.time() - t0
    S = args.ref_sample
    times = []
    for it in range(args.warmup + args.steps):
        r1, r2 = T.reads(cfg["read_seed"], S, L=cfg["L"], frag_mean=cfg["frag_mean"], frag_sd=cfg["frag_sd"], circ_frac=cfg["circ_frac"],
                         first=it * S)
        t0 = time.perf_counter()
        R = oracle.Run(otx, obsj, 2, r1, r2, nthreads=cores)
        dt = time.perf_counter() - t0
        R.close()
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    v = S / (ms * 1e-3)
    sample = "%d pairs per step of the same synthetic workload (fresh reads every step), mapping only (reads pre-parsed in memory)" % S
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/u64 integer", "data": "synthetic",
            "config": config_json(cfg, args),
            "details": {"reference": "restated algorithm (oracle/: C++ restatement of SACollectorCircall + Circall_wt/bsj per-pair bodies, std::thread pool over "
                                     "1000-pair jobs, open-addressing k-mer table), NOT the Circall binary, which cannot be built in this image",
                        "cores": cores, "oracle_index_set_up_s": round(t_index, 1), "product_library_loaded": False},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)
    return 0


def oracle_indices(cg, oracle, T, cores, gtx, gbsj):
    """Oracle indices for the cpu_baseline leg of the GPU arm: the (unique) suffix arrays are taken from the GPU builder and verified by
    the oracle before use (the reference arm sorts its own)."""
    otx = oracle.Index.from_text_with_sa(T.tx_text, gtx.suffix_array(), nthreads=cores)
    obsj = oracle.Index.from_text_with_sa(T.bsj_text, gbsj.suffix_array(), nthreads=cores)
    return otx, obsj


def _bind_to_gpu_socket(local):
    """Experiment switch (--numa-affinity, off by default): run this rank on the CPUs NVML reports as local to
    its GPU, so that the page-locked staging buffers it allocates afterwards sit on that socket.  Returns the CPU set or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception as e:                                   # an experiment: never fatal
        sys.stderr.write("numa affinity not set: %r\n" % (e,))
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json config number (1-based)")
    ap.add_argument("--genes", type=int, default=0, help="override the gene count (smaller index for quick runs)")
    ap.add_argument("--pairs", type=int, default=0, help="override pairs per GPU per step")
    ap.add_argument("--batch", type=int, default=1 << 25, help="pairs per library batch (the whole 20 M-pair step of config 2 is one batch: 4 Mi -> 197.9, 10 M -> 207.2, 20 M -> 211.3 M pairs/s; the latency tails of the warp / general kernels are per launch)")
    ap.add_argument("--cpu-sample", type=int, default=400_000, help="pairs in the cpu_baseline sample")
    ap.add_argument("--ref-sample", type=int, default=400_000, help="pairs per step of the reference arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg")
    ap.add_argument("--e2e-batch", type=int, default=1 << 20, help="pairs per submit of the end-to-end leg (smaller than --batch: the copy of the first batch and the read-back of the last are not overlapped, so short batches shorten the step)")
    ap.add_argument("--e2e-offsets", action="store_true", help="end-to-end leg through cg_session_submit (offset arrays copied with the reads) instead of cg_session_submit_fixed")
    ap.add_argument("--numa-affinity", action="store_true", help="experiment: bind the rank to the CPUs local to its GPU before the page-locked buffers are allocated")
    ap.add_argument("--e2e-no-lists", action="store_true", help="experiment: the end-to-end leg without the per-record result lists (counts only)")
    ap.add_argument("--resident-sessions", type=int, default=1, help="sessions the device-resident leg deals its batches to (each on its own stream: the short-list tail kernels of one overlap the thread tiers of the others); the step is cut into that many batches unless --batch is smaller")
    ap.add_argument("--sessions", type=int, default=4, help="sessions alternated by the end-to-end leg (copies of one overlap kernels of the others)")
    args = ap.parse_args()
    _claim_stdout()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    affinity, all_cpus = None, os.sched_getaffinity(0)
    if args.numa_affinity:
        affinity = _bind_to_gpu_socket(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import circall_b200 as cg
    import synth
    synth.build()
    cfg = workload(args)
    L, n = cfg["L"], cfg["n_pairs"]
    T = synth.Txome(cfg["tx_seed"], cfg["n_genes"])
    t0 = time.time()
    gtx = cg.Index.build(T.tx_text, device=local)
    gbsj = cg.Index.build(T.bsj_text, device=local)
    t_index = time.time() - t0
    # this rank's reads (weak scaling: every GPU maps its own n pairs)
    pin1, pin2 = cg.PinnedBuffer(n * L), cg.PinnedBuffer(n * L)
    r1, r2 = pin1.array.reshape(n, L), pin2.array.reshape(n, L)
    T.reads(cfg["read_seed"], n, L=L, frag_mean=cfg["frag_mean"], frag_sd=cfg["frag_sd"], circ_frac=cfg["circ_frac"], first=rank * n, out=(r1, r2))
    RS = max(1, args.resident_sessions)
    batch = min(args.batch, (n + RS - 1) // RS)
    off = np.arange(batch + 1, dtype=np.uint64) * np.uint64(L)
    pinoff = cg.PinnedBuffer(off.nbytes); pinoff.array[:] = off.view(np.uint8)
    d1, d2, doff = cg.DeviceBuffer(n * L + 64, local), cg.DeviceBuffer(n * L + 64, local), cg.DeviceBuffer(off.nbytes, local)
    d1.upload(r1); d2.upload(r2); doff.upload(off)
    batches = [(b0, min(batch, n - b0)) for b0 in range(0, n, batch)]

    S = cg.Session(gtx, gbsj, stage=0)
    RSS = [S] + [cg.Session(gtx, gbsj, stage=0) for _ in range(RS - 1)]
    accp, accn = S.accumulator()

    def acc_tensor(sess):
        p, nn = sess.accumulator()

        class _Acc:
            __cuda_array_interface__ = {"shape": (nn,), "typestr": "<i8", "data": (p, False), "version": 3}
        return torch.as_tensor(_Acc(), device=torch.device("cuda", local))
    acc_t = acc_tensor(S) if world > 1 else None
    acc_rs = [acc_tensor(x) for x in RSS[1:]] if world > 1 else []
    reduced = torch.empty(accn, dtype=torch.int64, device=torch.device("cuda", local)) if world > 1 else None

    def step_resident(stats=None):
        for i0 in range(0, len(batches), RS):
            grp = batches[i0:i0 + RS]
            for sess, (b0, nb) in zip(RSS, grp):
                sess.submit_device(d1.ptr + b0 * L, doff.ptr, d2.ptr + b0 * L, doff.ptr, nb, L)
            for sess, _ in zip(RSS, grp):
                B = sess.fetch(want_lists=False)
                if stats is not None:
                    stats.append(B)
        if world > 1:   # the one collective of the path: per-BSJ count vector + scalar counters (SURVEY §8 e)
            reduced.copy_(acc_t)
            for t in acc_rs:
                reduced.add_(t)
            dist.all_reduce(reduced)

    def bracket(fn, k):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(k):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t
        if world > 1:
            m = torch.tensor([dt], dtype=torch.float64, device=torch.device("cuda", local))
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
            dt = float(m.item())
        return dt

    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(local)
    clocks.start()
    stats = []
    l0 = sum(x.launches for x in RSS)
    dt = bracket(lambda: step_resident(stats), args.steps)
    launches = sum(x.launches for x in RSS) - l0
    clk = clocks.stop()
    ms_step = 1e3 * dt / args.steps
    value = world * n / (ms_step * 1e-3)
    kms = np.array([b.kernel_ms for b in stats])            # per launch: pack, wt, commit, bsj
    dev_ms = float(sum(b.device_ms for b in stats)) / args.steps
    k_avg = kms.mean(axis=0)
    n_exp = sum(b.n_export for b in stats) // args.steps
    n_junc = sum(b.n_junc for b in stats) // args.steps

    # ---- end to end through the C ABI from pinned host buffers, two sessions so copies overlap kernels
    e2e = None
    if not args.no_e2e:
        SS = RSS + [cg.Session(gtx, gbsj, stage=0) for _ in range(max(1, args.sessions) - len(RSS))]
        NS = len(SS)
        acc_all = [acc_t] + acc_rs + [acc_tensor(x) for x in SS[len(RSS):]] if world > 1 else []
        ebatch = min(args.e2e_batch, batch)
        # submit sizes ramp up (and down at the end): the copy of the very first submit and the read-back of the very last one
        # cannot overlap anything, so they are kept short
        ebatches, b0, size = [], 0, max(ebatch // 8, 1 << 16)
        while b0 < n:
            left = n - b0
            nb = min(size, left)
            if left - nb < ebatch // 4:          # the tail: two short submits instead of one long
                nb = left if left <= ebatch // 4 else left - ebatch // 8
            ebatches.append((b0, nb))
            b0 += nb
            size = min(2 * size, ebatch)
        h2d = 2 * n * L + (2 * sum(8 * (nb + 1) for _, nb in ebatches) if args.e2e_offsets else 0)
        d2h = [0]

        def step_host():
            pend = []
            tot = 0
            for i, (b0, nb) in enumerate(ebatches):
                s = SS[i % NS]
                if len(pend) == NS:
                    B = pend.pop(0).fetch(want_lists=not args.e2e_no_lists, materialize=False)
                    tot += B.d2h_bytes
                if args.e2e_offsets:
                    s.submit_raw(pin1.ptr + b0 * L, pinoff.ptr, pin2.ptr + b0 * L, pinoff.ptr, nb)
                else:                           # every read of the workload has the same length: cg_session_submit_fixed, no offsets on PCIe
                    s.submit_fixed_raw(pin1.ptr + b0 * L, pin2.ptr + b0 * L, nb, L)
                pend.append(s)
            for s in pend:
                B = s.fetch(want_lists=not args.e2e_no_lists, materialize=False)
                tot += B.d2h_bytes
            d2h[0] = tot
            if world > 1:            # every session's accumulator goes into the one all-reduce of the step
                reduced.copy_(acc_all[0])
                for t in acc_all[1:]:
                    reduced.add_(t)
                dist.all_reduce(reduced)
        for _ in range(2):
            step_host()
        dte = bracket(step_host, args.steps)
        e2e = {"value": world * n / (dte / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h[0]),
               "ms_per_step": 1e3 * dte / args.steps, "host_buffers": "pinned", "sessions": NS, "batch_pairs": ebatch,
               "call": "cg_session_submit" if args.e2e_offsets else "cg_session_submit_fixed"}
        # what the host-to-device path of this box gives: 1 GiB page-locked pieces, every rank copying at the same time (the concurrent
        # ceiling the end-to-end leg is measured against: at N = 8 the ranks share the host's memory and PCIe root complexes)
        try:
            hp = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
            dv = torch.empty(1 << 30, dtype=torch.uint8, device=torch.device("cuda", local))
            dv.copy_(hp, non_blocking=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(4):
                dv.copy_(hp, non_blocking=True)
            ev1.record()
            torch.cuda.synchronize()
            el = ev0.elapsed_time(ev1) * 1e-3
            if world > 1:                               # all ranks copy at once: the slowest rank's time bounds the aggregate
                t = torch.tensor([el], dtype=torch.float64, device=torch.device("cuda", local))
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                el = float(t.item())
            gbs_all = world * 4 * (1 << 30) / el / 1e9
            e2e["h2d_ceiling_gbs"] = gbs_all
            e2e["h2d_gbs"] = world * h2d / (dte / args.steps) / 1e9
            e2e["frac_of_h2d_ceiling"] = e2e["h2d_gbs"] / gbs_all
            e2e["h2d_ceiling_how"] = "4 x 1 GiB cudaMemcpyAsync from page-locked memory per rank, all ranks at once: total bytes / slowest rank's time"
            del hp, dv
        except Exception as ex:                                   # a measurement beside the metric: never fatal
            e2e["h2d_ceiling_gbs"] = None
            sys.stderr.write("h2d ceiling not measured: %r\n" % (ex,))
        if affinity:
            e2e["cpu_affinity"] = affinity
        for x in SS[len(RSS):]:
            x.close()
    if affinity:
        os.sched_setaffinity(0, all_cpus)                    # the cpu_baseline leg gets every host thread back

    # ---- cpu baseline + algorithmic bytes from the oracle's counters (rank 0, N = 1 only)
    cpu = None
    roof = None
    gather = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    if rank == 0:
        gather = cg.bench_gather(16 << 30, 64, 3, device=local)
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle
        oracle.build()
        cores = os.cpu_count() or 1
        t0 = time.time()
        otx, obsj = oracle_indices(cg, oracle, T, cores, gtx, gbsj)
        t_oidx = time.time() - t0
        ns = min(args.cpu_sample, n)
        t0 = time.perf_counter()
        R = oracle.Run(otx, obsj, 2, r1[:ns], r2[:ns], nthreads=cores)
        tc = time.perf_counter() - t0
        cw, cb = R.counters(0), R.counters(1)
        nj = R.bsj_junctions().shape[0]
        # parity of the benchmarked path on the sample: the FULL comparison of the GPU tests (export list, per-mate flags, junction
        # rows as a multiset, candidates, per-record hits, counters, equivalence classes + per-BSJ vector), not a few counters
        from tests.helpers import assert_batch_equals_oracle
        Sp = cg.Session(gtx, gbsj, stage=0, keep_mate_detail=True)
        Bp = Sp.run(r1[:ns], r2[:ns])
        cp, pb = Sp.counters()
        ok, why = True, None
        try:
            assert_batch_equals_oracle(Bp, cp, pb, R)
        except AssertionError as e:
            ok, why = False, str(e)
        Sp.close()
        cpu = {"value": ns / tc, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d pairs of the workload, wt + bsj on all host threads, mapping only (reads pre-parsed); oracle index set-up %.0f s not timed" % (ns, t_oidx),
               "parity_on_sample": bool(ok), "parity_means": "set equality with the oracle on the sample: export list, per-mate flags and hit counts, junction rows, "
               "candidates, per-record hits / split, 8 counters, equivalence classes and per-BSJ counts", "parity_first_difference": why}
        by_wt = algorithmic_bytes(cw, L, 0) / ns          # per pair
        by_bsj = algorithmic_bytes(cb, L, nj) / ns
        pairs_per_launch = n / len(batches)
        kern = {"k_wt": (by_wt, k_avg[1]), "k_bsj": (by_bsj, k_avg[3])}
        dom = max(kern, key=lambda k: kern[k][1])
        traffic = None
        try:   # ncu's DRAM bytes of one captured batch, scaled to this run's launch size (the capture records its own)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj[dom] * pairs_per_launch / tj.get("pairs_per_batch", 4_000_000)
        except Exception:
            pass
        ach = kern[dom][0] * pairs_per_launch / (kern[dom][1] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_pair": {"k_wt": by_wt, "k_bsj": by_bsj},
                "pairs_per_launch": pairs_per_launch, "avg_launch_ms": {"k_pack": k_avg[0], "k_wt": k_avg[1], "k_wt_commit+compaction": k_avg[2], "k_bsj": k_avg[3]},
                "random_sector_peak_gbs": gather, "frac_of_random_sector_peak": ach / gather if gather else None,
                "all_kernels": {k: {"achieved": v[0] * pairs_per_launch / (v[1] * 1e-3) / 1e9, "frac": v[0] * pairs_per_launch / (v[1] * 1e-3) / 1e9 / peak} for k, v in kern.items()},
                "bytes_source": "oracle work counters on the cpu_baseline sample (SURVEY.md §8 d formula), scaled to the launch size"}
    if rank == 0:
        ctr = S.counters(per_bsj=False)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/u64 integer", "data": "synthetic",
                "config": config_json(cfg, args),
                "details": {"n_tx": gtx.n_tx, "n_bsj": gbsj.n_tx, "tx_text_len": gtx.text_len, "bsj_text_len": gbsj.text_len, "batch_pairs": args.batch,
                            "index_build_s": round(t_index, 2), "index_device_gb": round((gtx.device_bytes + gbsj.device_bytes) / 1e9, 2),
                            "exported_records_per_step": int(n_exp), "junction_rows_per_step": int(n_junc)},
                "device_ms_per_step": dev_ms, "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu,
                "counters": {k: int(v) for k, v in ctr.items()}}
        _emit(line)
    for x in RSS:
        x.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
