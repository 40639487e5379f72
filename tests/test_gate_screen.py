"""The closed-form screen of the line gate (register.cu sym3_gate_screen) against the oracle's Jacobi eigenvalues
(oracle/liodom_oracle.cc sym3_eigenvalues, src/laser_odometry.cc:325-344), on the CPU: a NumPy restatement of the
screen must never DECIDE differently from `lambda2 > 3 lambda1` on the Jacobi values, on real neighbour sets (C1
data) and on synthetic near-threshold / degenerate scatters, and it must decide almost always.  On the GPU the same
check runs inside the kernel whenever per-edge outputs are requested (gate bit 2, test_gpu_register.py)."""
import numpy as np

import oracle
from conftest import get_sequence


def screen(c):
    """c: (n, 6) = c00, c01, c02, c11, c12, c22 -> +1 pass, 0 fail, -1 too close to call."""
    a00, a01, a02, a11, a12, a22 = c.T
    with np.errstate(all="ignore"):
        q = (a00 + a11 + a22) / 3.0
        b00, b11, b22 = a00 - q, a11 - q, a22 - q
        p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * (a01 * a01 + a02 * a02 + a12 * a12)
        p = np.sqrt(p2 / 6.0)
        ip = 1.0 / p
        c00, c11, c22, c01, c02, c12 = b00 * ip, b11 * ip, b22 * ip, a01 * ip, a02 * ip, a12 * ip
        r = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02))
        r = np.minimum(1.0, np.maximum(-1.0, r))
        phi = np.arccos(r) / 3.0
        l2 = q + 2.0 * p * np.cos(phi)
        l0 = q + 2.0 * p * np.cos(phi + 2.0943951023931953)
        l1 = 3.0 * q - l2 - l0
        dd, margin = l2 - 3.0 * l1, 1e-5 * l2
        out = np.where(dd > margin, 1, np.where(dd < -margin, 0, -1))
    return np.where(p2 > 0.0, out, -1)


def scatter(nn):
    """(n, 5, 3) float32 neighbours -> (n, 6) scatter entries in double, neighbour order (src/laser_odometry.cc:325-340)."""
    p = nn.astype(np.float64)
    m = p[:, 0]
    for r in range(1, 5):
        m = m + p[:, r]
    m = m / 5.0
    c = np.zeros((len(p), 6))
    for r in range(5):
        d = p[:, r] - m
        c[:, 0] += d[:, 0] * d[:, 0]; c[:, 1] += d[:, 0] * d[:, 1]; c[:, 2] += d[:, 0] * d[:, 2]
        c[:, 3] += d[:, 1] * d[:, 1]; c[:, 4] += d[:, 1] * d[:, 2]; c[:, 5] += d[:, 2] * d[:, 2]
    return c


def jacobi_gate(c):
    lib = oracle.lib()
    import ctypes
    out = np.empty(len(c), bool)
    w = np.empty(3)
    A = np.empty(9)
    for i, (a00, a01, a02, a11, a12, a22) in enumerate(c):
        A[:] = (a00, a01, a02, a01, a11, a12, a02, a12, a22)
        lib.orc_sym3_eigenvalues(A.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p))
        out[i] = w[2] > 3.0 * w[1]
    return out


def test_screen_on_c1_neighbour_sets():
    scans, gt = get_sequence("hdl64_small", 1000, 8)
    op = oracle.make_params(prev_frames=5)
    edges = [oracle.extract_scan(op, s)[0] for s in scans]
    decided = total = 0
    for f in range(3, 8):
        world = np.concatenate([oracle.transform(edges[g], gt[g]) for g in range(max(0, f - 5), f)])
        out = oracle.associate(edges[f], gt[f], world)
        ok = (out["gate"] & 1) != 0
        nn = world[out["knn_idx"][ok]][:, :, :3]
        s = screen(scatter(nn))
        exact = (out["gate"][ok] & 2) != 0
        called = s >= 0
        assert np.array_equal(s[called] == 1, exact[called])
        decided += int(called.sum()); total += len(s)
    assert total > 2000 and decided > 0.999 * total, (decided, total)


def test_screen_near_threshold_and_degenerate():
    rng = np.random.default_rng(5)
    n = 4000
    # random orthogonal bases, eigenvalues placed around lambda2 = 3 lambda1 at relative distances 1e-12 ... 1e-1
    Q = np.linalg.qr(rng.normal(size=(n, 3, 3)))[0]
    l1 = 10.0 ** rng.uniform(-6, 0, n)
    eps = np.concatenate([[0.0], 10.0 ** rng.uniform(-12, -1, n - 1)]) * rng.choice([-1.0, 1.0], n)
    l2 = 3.0 * l1 * (1.0 + eps)
    l0 = l1 * rng.uniform(0.0, 1.0, n)
    l0[::7] = l1[::7]          # double eigenvalue below
    A = np.einsum("nij,nj,nkj->nik", Q, np.stack([l0, l1, l2], 1), Q)
    c = np.stack([A[:, 0, 0], A[:, 0, 1], A[:, 0, 2], A[:, 1, 1], A[:, 1, 2], A[:, 2, 2]], 1)
    s = screen(c)
    exact = jacobi_gate(c)
    called = s >= 0
    assert np.array_equal(s[called] == 1, exact[called])
    assert np.all(s[np.abs(eps) < 1e-6] == -1)       # inside the margin: never called
    assert called[np.abs(eps) > 1e-3].all()          # far from it: always called
    # degenerate scatters: zero matrix, multiples of the identity, NaN
    z = np.zeros((3, 6)); z[1, [0, 3, 5]] = 2.5; z[2, 0] = np.nan
    sz = screen(z)
    assert sz[0] == -1 and sz[1] == -1 and sz[2] == -1
