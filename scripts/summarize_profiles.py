"""Turn the ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python scripts/summarize_profiles.py <tag> <launches.csv> <full.ncu-rep> [pairs per batch of the profiled run]

  profiles/<tag>_launches.txt      per-kernel launch count, total and share of the device time (launch list pass)
  profiles/<tag>_kernels.csv       headline metrics of every captured launch (ncu --set full)
  profiles/traffic.json            dram bytes (read + write) per launch of the dominant kernels, read by bench.py
"""
import csv
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    m = re.search(r"(k_\w+|cg\w*::\w+|\w+Kernel\w*|\w+)\s*(<[^(]*>)?\(", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def main():
    tag, launches, rep = sys.argv[1:4]
    pairs_per_batch = int(sys.argv[4]) if len(sys.argv) > 4 else 8_000_000      # of the profiled command (bench.py --pairs 8000000, one batch)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    # ---- launch list
    rows = [r for r in csv.reader(open(launches)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        t = float(r[vi].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        t *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        k = short(r[ki])
        tot[k][0] += 1
        tot[k][1] += t
    allt = sum(v[1] for v in tot.values())
    with open(os.path.join(ROOT, "profiles", tag + "_launches.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare shares)\n")
        f.write("# %-58s %8s %12s %8s\n" % ("kernel", "launches", "total ms", "share"))
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %8d %12.3f %7.1f%%\n" % (k[:60], n, t, 100 * t / allt))
        step = {k: v for k, v in tot.items() if re.search(r"k_pack|k_wt|k_bsj|k_thr|k_flag|k_cls_|k_exp|k_cand|k_acc|k_fixed|k_merge", k)}
        st = sum(v[1] for v in step.values())
        f.write("\n# kernels of the timed step only (index construction above is set-up)\n")
        for k, (n, t) in sorted(step.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %8d %12.3f %7.1f%%\n" % (k[:60], n, t, 100 * t / st))
    # ---- full capture
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u = rr[0], rr[1]
    stalls = [x for x in h if x.startswith("smsp__average_warps_issue_stalled") and x.endswith("per_issue_active.ratio")]
    traffic = {}
    with open(os.path.join(ROOT, "profiles", tag + "_kernels.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "value", "unit"])
        for r in rr[2:]:
            k = short(r[h.index("Kernel Name")])
            for m in KEEP + stalls:
                if m in h:
                    w.writerow([k, m, r[h.index(m)], u[h.index(m)]])
            rd, wr = r[h.index("dram__bytes_read.sum")], r[h.index("dram__bytes_write.sum")]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            b = float(rd) * scale[u[h.index("dram__bytes_read.sum")]] + float(wr) * scale[u[h.index("dram__bytes_write.sum")]]
            # the lockstep tier's collector is one template for both stages: k_thr_col<VOTE, BSJ>
            key = "k_wt" if ("k_wt" in k or re.search(r"k_thr_col<\w+, (0|false)>", k)) else "k_bsj" if ("k_bsj" in k or re.search(r"k_thr_col<\w+, (1|true)>", k)) else None
            if key and "k_wt_commit" not in k:
                traffic.setdefault(key, {}).setdefault(k, []).append(b)
            if "k_pack" in k:
                traffic.setdefault("_batches", []).append(1)
    # a stage (what bench.py brackets with events as k_wt / k_bsj) = all of its tiers: bytes of every captured launch of the stage, per batch
    nbat = max(1, len(traffic.pop("_batches", [])))
    tj = {key: sum(sum(v) for v in d.values()) / nbat for key, d in traffic.items()}
    tj["pairs_per_batch"] = pairs_per_batch
    tj["_note"] = "dram__bytes_read.sum + dram__bytes_write.sum per batch, summed over the kernels (tiers) of the stage, ncu --set full, %s" % os.path.basename(rep)
    json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(open(os.path.join(ROOT, "profiles", tag + "_launches.txt")).read())
    print(json.dumps(tj, indent=1))


if __name__ == "__main__":
    main()
