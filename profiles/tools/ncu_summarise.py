#!/usr/bin/env python
"""Turn the ncu outputs of a gpurun call into the small tracked summaries under profiles/:

    python profiles/tools/ncu_summarise.py launches gpurun_out/launches_X.csv profiles/X_launches_summary.csv
    python profiles/tools/ncu_summarise.py raw gpurun_out/prof_X.ncu-rep profiles/X_k_search_ncu_summary.csv

`launches`: per-kernel totals of the `--metrics gpu__time_duration.sum` launch list (shares, not
absolutes: launches are cold-cache and serialised under ncu).  `raw`: selected metrics of the
`--set full` capture, one column per kernel (first captured launch of each)."""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "inst_executed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct"]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    mu = hdr.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[1:]:
        v = float(r[mv].replace(",", ""))
        v = v / 1000.0 if r[mu] in ("ns", "nsecond") else (v * 1000.0 if r[mu] in ("ms", "msecond") else v)
        name = r[kn][:70]
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + v)
    total = sum(t for _, t in tot.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_us,us_per_launch,share_of_all_launches\n")
        for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{name}",{n},{t:.1f},{t / n:.1f},{t / total:.4f}\n')


def raw(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    first = OrderedDict()
    for r in data:
        first.setdefault(r[kn], r)
    cols = [h for h in hdr if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"))]
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + list(first.keys()))
        for c in cols:
            i = hdr.index(c)
            w.writerow([c, units[i]] + [r[i] for r in first.values()])


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
