#!/usr/bin/env python
"""Per-kernel SASS summary of the built objects (instruction count, registers, and the mnemonics that show
how each kernel maps to sm_100a: UBLKCP = cp.async.bulk (TMA), SYNCS = mbarrier, CREDUX/REDUX = warp
reductions, MATCH = match_any, ATOM/ATOMG/RED = global atomics, F2FP/D* = FP64 pipeline, UCGABAR/CCTL = cluster).

    python tools/sass_evidence.py > profiles/sass_evidence_rNN.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UBLKCP", "SYNCS", "CREDUX", "REDUX", "MATCH", "VOTE", "SHFL", "ATOMG", "ATOM", "RED", "LDG", "STG", "LDS", "STS", "LDL", "STL",
        "DADD", "DMUL", "DFMA", "MUFU", "FFMA", "UCGABAR", "BAR", "LDGDEPBAR", "PREFETCH"]
for obj in sorted(glob.glob(os.path.join(ROOT, "liodom_b200", "csrc", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res))
    print("== %s" % os.path.basename(obj))
    name, cnt = None, None
    out = []
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if name:
                out.append((name, cnt))
            name, cnt = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and name:
            op = m.group(1)
            cnt["_total"] += 1
            base = op.split(".")[0]
            if base in KEYS:
                cnt[base] += 1
    if name:
        out.append((name, cnt))
    for name, c in out:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(anonymous namespace\)::", "", dem).split("(")[0] or name
        print("%-44s %5d instr  %3s regs  %s" % (dem, c["_total"], regs.get(name, "?"),
                                               " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))
