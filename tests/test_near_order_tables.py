"""The constant tables of the nearest-first cell order in liodom_b200/csrc/register.cu (kNearMask on the canonical
occupancy word, kNearOff, the three reflection mask triples of occ_canonical), parsed from the source and checked
against the rule they encode: slot i visits, per axis, the own cell (0), the neighbour behind the nearer face (near) or
the one behind the farther face (far), slots sorted by the sum of per-axis weights 0 / 1 / 4; raw occupancy bit
((dz+1)*3 + (dy+1))*3 + (dx+1)."""
import itertools
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "liodom_b200", "csrc", "register.cu")).read()


def _table(name):
    m = re.search(name + r"\[27\]\s*=\s*\{(.*?)\};", SRC, re.S)
    return [int(t, 0) for t in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(1))]


def _canonical(occ, nx, ny, nz):
    m = re.search(r"unsigned occ_canonical\(.*?\{(.*?)return occ;", SRC, re.S).group(1)
    rows = re.findall(r"if \(n([xyz]) > 0\) occ = \(\(occ & (0x[0-9a-f]+)u\) << (\d+)\) \| \(\(occ & (0x[0-9a-f]+)u\) >> (\d+)\) \| \(occ & (0x[0-9a-f]+)u\);", m)
    assert [r[0] for r in rows] == ["x", "y", "z"]
    for (ax, lo, sh1, hi, sh2, mid), n in zip(rows, (nx, ny, nz)):
        if n > 0:
            occ = ((occ & int(lo, 16)) << int(sh1)) | ((occ & int(hi, 16)) >> int(sh2)) | (occ & int(mid, 16))
    return occ


def test_tables_match_the_rule():
    mask, off = _table("kNearMask"), _table("kNearOff")
    assert len(mask) == 27 and len(off) == 27 and sorted(mask) == [1 << b for b in range(27)]
    weight = {0: 0, -1: 1, 1: 4}   # canonical offset: 0 own, -1 near, +1 far

    def sx(v, sh):
        x = (v << sh) & 0xffffffff
        return ((x - (1 << 32)) if x & 0x80000000 else x) >> 30
    canon = [(sx(v, 30), sx(v, 28), sx(v, 26)) for v in off]
    assert canon[0] == (0, 0, 0) and len(set(canon)) == 27
    sums = [sum(weight[c] for c in t) for t in canon]
    assert sums == sorted(sums)                                     # nearest first
    for m, (cx, cy, cz) in zip(mask, canon):
        assert m == 1 << (((cz + 1) * 3 + (cy + 1)) * 3 + (cx + 1))  # the slot's bit on the canonical word
    rng = np.random.default_rng(2)
    for nx, ny, nz in itertools.product((-1, 1), repeat=3):
        for occ in [int(v) for v in rng.integers(0, 1 << 27, 200)]:
            oc = _canonical(occ, nx, ny, nz)
            for m, (cx, cy, cz) in zip(mask, canon):
                dx, dy, dz = -cx * nx, -cy * ny, -cz * nz          # near_offsets(): canonical offset times the near direction
                raw = 1 << (((dz + 1) * 3 + (dy + 1)) * 3 + (dx + 1))
                assert bool(occ & raw) == bool(oc & m)
