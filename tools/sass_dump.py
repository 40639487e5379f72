#!/usr/bin/env python
"""Full SASS listing of every kernel of the built objects, one file per kernel under profiles/sass/
(cuobjdump -sass; sm_100a).  The mnemonic summary next to it is tools/sass_evidence.py.

    python tools/sass_dump.py
"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "sass")
os.makedirs(OUT, exist_ok=True)
for f in glob.glob(os.path.join(OUT, "*.sass")):
    os.remove(f)
index = []
for obj in sorted(glob.glob(os.path.join(ROOT, "liodom_b200", "csrc", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
    usage = dict(re.findall(r"Function (\S+):\s*\n\s*(REG:\d+[^\n]*)", res))
    cur, name = [], None

    def flush():
        if not name:
            return
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*$", "", dem.replace("(anonymous namespace)::", "")).replace("liodom::", "").replace("void ", "")
        short = re.sub(r"[^A-Za-z0-9_<>,]", "", short).replace("<", "_").replace(">", "").replace(",", "_")
        path = os.path.join(OUT, "%s__%s.sass" % (os.path.basename(obj)[:-2], short))
        body = [l for l in cur if l.strip()]
        ninstr = sum(1 for l in body if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l))
        with open(path, "w") as fh:
            fh.write("// %s\n// %s (%s), sm_100a, %d instructions, %s\n" % (dem, name, os.path.basename(obj), ninstr, usage.get(name, "")))
            fh.write("\n".join(body) + "\n")
        index.append((os.path.basename(path), ninstr, usage.get(name, "")))

    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            flush()
            name, cur = m.group(1), []
            continue
        if name:
            # drop the encoding column: keep address + instruction
            cur.append(re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", ln.rstrip()))
    flush()
with open(os.path.join(OUT, "INDEX.txt"), "w") as fh:
    fh.write("# cuobjdump -sass of liodom_b200/csrc/*.o (nvcc -gencode arch=compute_100a,code=sm_100a), one file per kernel\n")
    for n, k, u in sorted(index):
        fh.write("%-70s %6d instr  %s\n" % (n, k, u))
print(len(index), "kernels")
