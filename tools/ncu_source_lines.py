#!/usr/bin/env python
"""Per-source-line hot spots from an ncu report captured with --import-source on (-lineinfo builds).

  python tools/ncu_source_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import subprocess
import sys


def main(path, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    fname = ""
    lines = []
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "":
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")

        def num(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        lines.append((num(r[si]), num(r[ii]), num(r[ti]), fname, ln, r[1].strip()))
    tot_s = sum(l[0] for l in lines) or 1.0
    tot_i = sum(l[1] for l in lines) or 1.0
    print("# %s: total samples %d, warp instructions %d, avg active threads %.1f"
          % (path, tot_s, tot_i, sum(l[2] for l in lines) / tot_i))
    print("%7s %7s %6s  %s" % ("samp%", "inst%", "thr", "file:line  source"))
    for s, i, t, f, ln, src in sorted(lines, key=lambda x: -x[0])[:top]:
        print("%7.2f %7.2f %6.1f  %s:%d  %s" % (100 * s / tot_s, 100 * i / tot_i, t / i if i else 0, f, ln, src[:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
