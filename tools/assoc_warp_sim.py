#!/usr/bin/env python
"""Warp-level replay of the association's candidate loop on C1 data (CPU only: oracle + NumPy).

For every edge of one steady-state frame the exact search is replayed (own 0.5 m cell first, then the neighbour cells
in the kernel's nearest-first order under the progressively tightened bound), and 32 Morton-consecutive edges are put
in one warp as k_edge_order does.  Printed per warp, in units of "iterations of 4 candidates" (the unrolled loop of
knn_scan_bucket):

  lockstep-slot   what k_associate<1> executes: all lanes step through the 26 neighbour slots together
  per-lane pop    every lane walks ITS occupied cells (j-th occupied / j-th unpruned cell of every lane together)
  flattened       every lane walks all its candidates back to back (max over lanes of the per-lane total)
  pooled          the warp's candidates divided evenly over the 32 lanes (sum / 32)

and, for the warp-pooled kernel (k_associate_pool: own cell + nearest cells in place until five candidates are in hand,
the rest listed and scanned by the whole warp under the static bound), the in-place iterations, the pooled iterations and
the number of pooled candidates that pass their owner's bound.  DESIGN.md section 5 quotes these numbers.

    python tools/assoc_warp_sim.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, oracle
from liodom_b200 import synth
from collections import defaultdict

scans, gt = synth.sequence("hdl64", 1000, 18)
op = oracle.make_params(prev_frames=15)
odo = oracle.Odometer(op)
edges = [oracle.extract_scan(op, s)[0] for s in scans]
for f in range(17):
    odo.process(edges[f])
win, _ = odo.window(); W = win[:, :3].astype(np.float32)
o, pv = odo.get_pose()
pred = o @ (np.linalg.inv(pv) @ o)
q = (edges[17][:, :3].astype(np.float64) @ pred[:3, :3].T + pred[:3, 3]).astype(np.float32)
cell = 0.5
wc = np.floor(W / cell).astype(np.int64); qc = np.floor(q / cell).astype(np.int64)
bucket = defaultdict(list)
for i, c in enumerate(map(tuple, wc)):
    bucket[c].append(i)
bucket = {k: np.array(v) for k, v in bucket.items()}


def spread(v):
    v = v & 0x3ff; v = (v | (v << 16)) & 0x030000ff; v = (v | (v << 8)) & 0x0300f00f
    v = (v | (v << 4)) & 0x030c30c3; v = (v | (v << 2)) & 0x09249249
    return v


mk = spread(qc[:, 0].astype(np.uint32)) | (spread(qc[:, 1].astype(np.uint32)) << 1) | (spread(qc[:, 2].astype(np.uint32)) << 2)
mo = np.argsort(mk, kind='stable')
# kNearOrder of register.cu: per axis 0 = own, 1 = near side, 2 = far side
near_codes = [0x00, 0x01, 0x04, 0x10, 0x05, 0x11, 0x14, 0x15, 0x02, 0x08, 0x20, 0x06, 0x09, 0x12, 0x18, 0x21, 0x24,
              0x16, 0x19, 0x25, 0x0a, 0x22, 0x28, 0x1a, 0x26, 0x29, 0x2a]
it4 = lambda n: (n + 3) // 4


def replay(i, pooled):
    """-> (own count, [(slot, scanned count or 0 if pruned)], listed counts, pooled candidates within the bound)"""
    c = qc[i]; qq = q[i]
    n = [(-1 if (qq[a] - cell * c[a]) < 0.5 * cell else 1) for a in range(3)]
    best = np.zeros(0, np.float32); k5 = np.inf
    own = bucket.get(tuple(c)); nown = 0

    def offer(b):
        nonlocal best, k5
        d2 = ((W[b] - qq) ** 2).sum(1)
        best = np.sort(np.concatenate([best, d2[d2 < 1.0]]))[:5]
        if len(best) == 5: k5 = best[4]
        return d2
    if own is not None:
        nown = len(own); offer(own)
    sc, listed, surv = [], [], 0
    for r, code in enumerate(near_codes[1:], 1):
        ax, ay, az = code & 3, (code >> 2) & 3, code >> 4
        d = [0 if a == 0 else (n[j] if a == 1 else -n[j]) for j, a in enumerate((ax, ay, az))]
        cc = (c[0] + d[0], c[1] + d[1], c[2] + d[2]); b = bucket.get(cc)
        if b is None: continue
        lo = np.array(cc) * cell; hi = lo + cell
        g = np.maximum(0, np.maximum(lo - qq, qq - hi)); dm = (g.astype(np.float32) ** 2).sum()
        if dm >= 1.0 or dm > k5: sc.append((r, 0)); continue
        if pooled and len(best) == 5:   # bound in hand: list, scanned later under the static bound
            d2 = ((W[b] - qq) ** 2).sum(1)
            listed.append(len(b)); surv += int(((d2 < 1.0) & (d2 <= k5)).sum())
        else:
            sc.append((r, len(b))); offer(b)
    return nown, sc, listed, surv


nw = len(q) // 32
res = [replay(i, False) for i in range(len(q))]
tot = sum(r[0] + sum(c for _, c in r[1]) for r in res)
print("thread-per-edge: %.1f candidates per edge (own cell %.1f)" % (tot / len(res), np.mean([r[0] for r in res])))
A = B = B2 = C = D = OWN = 0
for w in range(nw):
    lanes = [res[j] for j in mo[w * 32:(w + 1) * 32]]
    own = max(it4(l[0]) for l in lanes); OWN += own
    A += own + sum(max([it4(c) for l in lanes for rr, c in l[1] if rr == r] or [0]) for r in range(1, 27))
    m = max(len(l[1]) for l in lanes)
    B += own + sum(max([it4(l[1][j][1]) for l in lanes if len(l[1]) > j] or [0]) for j in range(m))
    un = [[c for _, c in l[1] if c > 0] for l in lanes]
    m = max(len(u) for u in un)
    B2 += own + sum(max([it4(u[j]) for u in un if len(u) > j] or [0]) for j in range(m))
    C += max(it4(l[0]) + sum(it4(c) for _, c in l[1]) for l in lanes)
    D += sum(it4(l[0]) + sum(it4(c) for _, c in l[1]) for l in lanes) / 32
print("iterations of 4 candidates per warp: own cells %.1f | lockstep-slot %.1f | per-lane pop %.1f (unpruned only %.1f) | "
      "flattened %.1f | pooled %.1f" % (OWN / nw, A / nw, B / nw, B2 / nw, C / nw, D / nw))

res = [replay(i, True) for i in range(len(q))]
S = np.array([r[3] for r in res]); L = np.array([sum(r[2]) for r in res]); NS = np.array([len(r[2]) for r in res])
P = np.array([r[0] + sum(c for _, c in r[1]) for r in res])
print("warp-pooled kernel: in place %.1f candidates per edge, listed %.1f in %.2f buckets (p99 %d, max %d); within the "
      "bound %.2f per edge (p90 %d, p99 %d, max %d)" % (P.mean(), L.mean(), NS.mean(), np.percentile(NS, 99), NS.max(),
                                                     S.mean(), *np.percentile(S, [90, 99]), S.max()))
Sm = S[mo][:nw * 32].reshape(nw, 32); Lm = L[mo][:nw * 32].reshape(nw, 32)
pre = 0
for w in range(nw):
    lanes = [res[j] for j in mo[w * 32:(w + 1) * 32]]
    pre += max(it4(l[0]) for l in lanes)
    pre += sum(max([it4(c) for l in lanes for rr, c in l[1] if rr == r] or [0]) for r in range(1, 27))
ws = Sm.sum(1)
print("per warp: in-place iterations %.1f, pooled iterations %.1f, candidates within the bound %.1f (largest per lane %.1f; "
      "p99 of the warp sum %d, %.1f %% of the warps above 224)" % (pre / nw, (Lm.sum(1) / 128).mean(), ws.mean(), Sm.max(1).mean(),
                                                                 np.percentile(ws, 99), 100 * (ws > 224).mean()))
