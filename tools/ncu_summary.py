#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/launches_rNN.txt
  python tools/ncu_summary.py kernel gpurun_out/prof.ncu-rep   > profiles/kernel_rNN.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("%-44s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-44s %6d %12.1f %10.1f %7.3f" % (k, a[0], a[1], a[1] / a[0], a[1] / tot))
    print("%-44s %6d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print("# ncu --set full --clock-control none: %s" % path)
    print("# kernels: %s" % ", ".join(sorted(set(r[ki].split("(")[0] for r in data))))
    for w in KEYS + sorted(h for h in hdr if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio")):
        if w in hdr:
            i = hdr.index(w)
            print("%-86s %-16s %s" % (w, units[i], " ".join(r[i] for r in data)))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
