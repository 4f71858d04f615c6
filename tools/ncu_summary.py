#!/usr/bin/env python
"""Summarises ncu outputs into small text files for profiles/ (run here, on the CPU box).
  python tools/ncu_summary.py launches <launches.csv> <out.txt>     # per-kernel totals and shares of the run
  python tools/ncu_summary.py full <prof.ncu-rep> <out.txt>         # key metrics of every captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(d["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1.0)
            a = agg.setdefault(d["Kernel Name"][:90], [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: per-kernel totals (cold-cache, serialised)\n")
        f.write("%-92s %6s %12s %7s\n" % ("kernel", "n", "total ms", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-92s %6d %12.3f %6.2f%%\n" % (k, v[0], v[1], 100 * v[1] / tot))
        f.write("total %.3f ms over %d launches\n" % (tot, sum(v[0] for v in agg.values())))


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none (raw page), selected metrics per captured launch\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("== %s\n" % d.get("Kernel Name", "?"))
            for k in KEYS:
                if k in d:
                    f.write("  %-82s %s %s\n" % (k, d[k], units[hdr.index(k)]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
