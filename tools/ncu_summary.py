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


STEP_KEYS = [("gpu__time_duration.sum", "time_us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
             ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
             ("lts__t_sector_hit_rate.pct", "l2_hit%"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
             ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
             ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("smsp__inst_executed.sum", "warp_inst")]


def step(path):
    """One row per captured launch: the whole step at a glance (every kernel of liodom_scan_batch)."""
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    keys = [(k, n) for k, n in STEP_KEYS if k in hdr]
    print("# ncu --set full --clock-control none: %s (cold cache, serialised: compare shares, not absolutes)" % path)
    print("%-20s" % "kernel" + " ".join("%11s" % n for _, n in keys))
    tot = 0.0
    for r in data:
        name = r[ki].split("(")[0].replace("void ", "")
        vals = []
        for k, _ in keys:
            v = r[hdr.index(k)].replace(",", "")
            try:
                f = float(v)
                if k == "gpu__time_duration.sum":
                    f = f / 1e3 if units[hdr.index(k)] in ("ns", "nsecond") else f
                    tot += f
                if units[hdr.index(k)] == "byte":
                    f /= 1e6
                if units[hdr.index(k)] == "Kbyte":
                    f /= 1e3
                if units[hdr.index(k)] == "Gbyte":
                    f *= 1e3
                vals.append("%11.2f" % f if f < 1e7 else "%11.4g" % f)
            except ValueError:
                vals.append("%11s" % v[:11])
        print("%-20s" % name[:20] + " ".join(vals))
    print("# total %.1f us over %d launches" % (tot, len(data)))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "step": step}[sys.argv[1]](sys.argv[2])
