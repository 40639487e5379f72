#!/usr/bin/env python
"""Where a kernel's warp instructions are, in program order: the SASS listing of an ncu report captured with
--import-source on, cut into blocks of N instructions, with executed warp instructions per warp, the average number of
active threads and the dominant opcodes per block; plus the position of every warp-level primitive (SHFL, VOTE, REDUX,
ATOMS, MATCH), which mark the boundaries between a kernel's phases.

    python tools/ncu_sass_blocks.py report.ncu-rep <warps per launch> [block size] [kernel index]
"""
import csv, subprocess, sys


def main(path, warps, block=50, which=0):
    raw = subprocess.run(["ncu", "-i", path, "--launch-skip", str(which), "--launch-count", "1", "--page", "source", "--print-source", "sass", "--csv"],
                         capture_output=True, text=True).stdout
    kernels, cur, hdr = [], None, None
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif r[0] == "Address":
            hdr = r
        elif cur is not None and hdr is not None and len(r) == len(hdr):
            cur["rows"].append(r)
    k = kernels[0]
    ia, ii, it = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    data = [(r[ia].strip(), float(r[ii]) / warps, float(r[it]) / warps) for r in k["rows"]]
    tot = sum(d[1] for d in data)
    print("# %s\n# %d SASS instructions, %.0f executed warp instructions per warp (%d warps), %.1f active threads"
          % (k["name"], len(data), tot, warps, sum(d[2] for d in data) / max(tot, 1e-9)))
    print("# sass range   executed/warp  cumulative  threads  max-exec  dominant opcodes")
    acc = 0.0
    for b in range(0, len(data), block):
        blk = data[b:b + block]
        n = sum(d[1] for d in blk); t = sum(d[2] for d in blk); acc += n
        if n < 0.5:
            continue
        ops = {}
        for d in blk:
            tok = d[0].split()
            op = (tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]).split(".")[0]
            ops[op] = ops.get(op, 0.0) + d[1]
        top = " ".join("%s:%.0f" % kv for kv in sorted(ops.items(), key=lambda x: -x[1])[:5])
        print("%5d-%-5d %12.1f %11.0f %8.1f %9.1f  %s" % (b, b + block, n, acc, t / max(n, 1e-9), max(d[1] for d in blk), top))
    print("# warp-level primitives (phase markers)")
    acc = 0.0
    for i, (s, n, t) in enumerate(data):
        acc += n
        if n > 0 and any(m in s for m in ("SHFL", "VOTE", "REDUX", "ATOMS", "MATCH")):
            print("%5d  %-48s executed/warp %6.2f  threads %4.1f  cumulative %6.0f" % (i, s[:48], n, t / max(n, 1e-9), acc))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 50, int(sys.argv[4]) if len(sys.argv) > 4 else 0)
