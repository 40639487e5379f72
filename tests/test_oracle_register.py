"""CPU tests: the C++ oracle of the LaserOdometer / Map path against independent NumPy/SciPy
re-derivations and the invariants read off the reference."""
import os

import numpy as np
import pytest

import oracle
import np_ref
from conftest import get_sequence, pose_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _edges_of(p, scans):
    out = []
    for s in scans:
        sp = oracle.split(p, s)
        out.append(oracle.extract(p, sp["rings"], sp["offsets"])["edges"])
    return out


def test_transform_double_math_float_store():
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(1000, 4)).astype(np.float32) * 30
    from scipy.spatial.transform import Rotation
    T = np.eye(4)
    T[:3, :3] = Rotation.from_rotvec([0.1, -0.2, 0.7]).as_matrix()
    T[:3, 3] = [12.3, -4.5, 0.67]
    assert np.array_equal(oracle.transform(pts, T).view(np.uint32), np_ref.transform(pts, T).view(np.uint32))


def test_knn_kdtree_equals_bruteforce_equals_numpy():
    g = np.load(os.path.join(GOLD, "register_hdl64_small.npz"))
    window, edges, T = g["window"], g["edges"], g["pose"]
    q = oracle.transform(edges, T)
    i0, d0, t0 = oracle.knn5(window, q, 0)
    i1, d1, t1 = oracle.knn5(window, q, 1)
    assert np.array_equal(i0, i1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32)) and np.array_equal(t0, t1)
    ni, nd = np_ref.knn5_bruteforce(window, q[:300])
    assert np.array_equal(ni, i0[:300]) and np.array_equal(nd.view(np.uint32), d0[:300].view(np.uint32))
    assert t0.mean() < 0.01   # the generator produces (almost) no exact distance ties


def test_knn_tie_flag_and_small_maps():
    m = np.zeros((6, 4), np.float32)
    m[:, 0] = [1, -1, 2, -2, 3, -3]                # symmetric around the query: d2 ties everywhere
    q = np.zeros((1, 4), np.float32)
    idx, d2, tie = oracle.knn5(m, q, 0)
    assert tie[0] == 1 and list(idx[0]) == [0, 1, 2, 3, 4]   # ties resolved by index
    idx, d2, tie = oracle.knn5(m[:3], q, 1)                    # fewer than 5 points
    assert list(idx[0][3:]) == [-1, -1] and np.isinf(d2[0][3:]).all()


def test_gates_and_line_through_first_two_neighbours():
    """d2[4] < 1.0 and lambda2 > 3 lambda1 (src/laser_odometry.cc:324,344); the line is through
    NN[0], NN[1], not the principal direction (:351-359)."""
    g = np.load(os.path.join(GOLD, "register_hdl64_small.npz"))
    window, edges, T = g["window"], g["edges"], g["pose"]
    a = oracle.associate(edges, T, window, knn_method=0)
    for k in ("knn_idx", "gate", "tie"):
        assert np.array_equal(a[k], g[k])
    assert np.array_equal(a["knn_d2"].view(np.uint32), g["knn_d2"].view(np.uint32))
    assert np.array_equal(a["eig"].view(np.uint64), g["eig"].view(np.uint64))
    g1 = (a["gate"] & 1) == 1
    assert np.array_equal(g1, a["knn_d2"][:, 4] < 1.0)
    chk = np.nonzero(g1)[0][:400]
    for i in chk:
        nn = window[a["knn_idx"][i], :3].astype(np.float64)
        c = nn.sum(0) / 5.0
        ev = np.linalg.eigvalsh((nn - c).T @ (nn - c))
        assert np.allclose(ev, a["eig"][i], rtol=1e-9, atol=1e-12)
        margin = abs(ev[2] - 3 * ev[1])
        if margin > 1e-9 * ev[2]:
            assert bool(a["gate"][i] & 2) == bool(ev[2] > 3 * ev[1])
    sel = (a["gate"] & 2) == 2
    cab = np.concatenate([edges[sel][:, :3], window[a["knn_idx"][sel][:, 0]][:, :3], window[a["knn_idx"][sel][:, 1]][:, :3]], 1)
    assert np.array_equal(cab.astype(np.float64), g["cab"])   # curr_point is the UN-transformed edge (:347-349)


def test_factor_residual_jacobian_and_weight():
    """Point2LineFactor (include/liodom/factors.hpp:71-105): residual vs NumPy, autodiff
    Jacobian vs central differences through EigenQuaternionParameterization::Plus, and the
    weight quirk w = 1.01 - (||(c - t)_xy|| - min)/(max - min) incl. its t-derivative."""
    rng = np.random.default_rng(1)
    from scipy.spatial.transform import Rotation
    for trial in range(20):
        c = rng.normal(size=3) * 20
        a = c + rng.normal(size=3) * 0.3
        b = a + rng.normal(size=3)
        q = Rotation.from_rotvec(rng.normal(size=3) * 0.2).as_quat()
        t = rng.normal(size=3) * (100 if trial % 2 else 1)        # far from the origin: w < 0
        r, J = oracle.factor(c, a, b, q, t)
        assert np.allclose(r, np_ref.point2line_residual(c, a, b, q, t), rtol=1e-12, atol=1e-12)
        h = 1e-6
        Jn = np.zeros((3, 6))
        for k in range(6):
            d = np.zeros(6)
            d[k] = h
            rp = np_ref.point2line_residual(c, a, b, np_ref.quat_plus(q, d[:3]), t + d[3:])
            rm = np_ref.point2line_residual(c, a, b, np_ref.quat_plus(q, -d[:3]), t - d[3:])
            Jn[:, k] = (rp - rm) / (2 * h)
        assert np.allclose(J, Jn, rtol=1e-5, atol=1e-6), (trial, np.abs(J - Jn).max())
    # weight: point 10 m in front of the sensor, translation at the origin -> w = 1.01 - 7/72
    r, _ = oracle.factor([10, 0, 0], [10, 1, 0], [10, 1, 1], [0, 0, 0, 1], [0, 0, 0])
    assert np.isclose(np.linalg.norm(r), (1.01 - 7.0 / 72.0) * 1.0)
    # the same geometry 200 m away: c stays in the sensor frame, t is the world translation
    r2, _ = oracle.factor([10, 0, 0], [210, 1, 0], [210, 1, 1], [0, 0, 0, 1], [200, 0, 0])
    assert np.isclose(np.linalg.norm(r2), abs(1.01 - (190.0 - 3.0) / 72.0))


def test_quaternion_plus_convention():
    """x_plus = dq (x) x with dq = (sin|d|/|d| d, cos|d|): NOT the half-angle (App. A.5)."""
    from scipy.spatial.transform import Rotation
    q = np.array([0, 0, 0, 1.0])
    d = np.array([0, 0, 0.1])
    qp = np_ref.quat_plus(q, d)
    assert np.isclose(Rotation.from_quat(qp).magnitude(), 0.2)   # delta 0.1 -> rotation by 0.2 rad
    # the oracle's LM uses the same Plus: a pure-rotation problem converges to the right rotation
    rng = np.random.default_rng(2)
    R = Rotation.from_rotvec([0.0, 0.0, 0.05])
    c = rng.normal(size=(200, 3)) * 10 + [20, 0, 0]
    lp = R.apply(c)
    dirs = rng.normal(size=(200, 3))
    cab = np.concatenate([c, lp + 0.3 * dirs, lp - 0.4 * dirs], 1)
    qo, to, s = oracle.solve(cab, [0, 0, 0, 1], [0, 0, 0])
    assert np.abs(to).max() < 1e-6 and Rotation.from_quat(qo / np.linalg.norm(qo)).inv().__mul__(R).magnitude() < 1e-6


def test_lm_against_independent_minimiser():
    """The restated ceres::Solve reduces the Huber cost like an independent solver does, QR and
    Cholesky linear solvers agree, and the controller honours max_num_iterations = 4."""
    g = np.load(os.path.join(GOLD, "register_hdl64_small.npz"))
    cab, q0, t0 = g["cab"], g["q0"], g["t0"]
    qa, ta, sa = oracle.solve(cab, q0, t0, linear_solver=0)
    qb, tb, sb = oracle.solve(cab, q0, t0, linear_solver=1)
    assert (sa.iterations, sa.termination) == (sb.iterations, sb.termination)
    assert np.abs(ta - tb).max() < 1e-9 and np.abs(qa - qb).max() < 1e-10
    assert np.array_equal(qa, g["q1"]) and np.array_equal(ta, g["t1"])        # golden
    assert [sa.iterations, sa.successful_steps, sa.termination, sa.num_residual_blocks] == list(g["summary"])
    assert sa.iterations <= 4 and sa.jac_evals == sa.successful_steps + 1
    c0 = np_ref.huber_cost(cab, q0, t0)
    c1 = np_ref.huber_cost(cab, qa, ta)
    assert np.isclose(c0, sa.initial_cost, rtol=1e-10) and np.isclose(c1, sa.final_cost, rtol=1e-10)
    assert c1 < c0
    # independent: scipy on the 6-dof tangent space with the same robust loss
    from scipy.optimize import least_squares

    def fun(x):
        q = np_ref.quat_plus(q0, x[:3])
        res = []
        for row in cab:
            r = np_ref.point2line_residual(row[0:3], row[3:6], row[6:9], q, t0 + x[3:])
            s = float(r @ r)
            res.append(np.sqrt(s if s <= 0.04 else 0.4 * np.sqrt(s) - 0.04))
        return np.array(res)
    sub = slice(0, len(cab), max(1, len(cab) // 300))
    cab_s = cab[sub]
    qs, ts, ss = oracle.solve(cab_s, q0, t0)
    cab = cab_s
    best = least_squares(fun, np.zeros(6), method="lm", max_nfev=200)
    c_best = 0.5 * float(best.fun @ best.fun)
    assert ss.final_cost <= ss.initial_cost
    # four LM iterations get within a few percent of the converged optimum on this problem
    assert ss.final_cost <= c_best * 1.05 + 1e-9


def test_local_map_manager_window():
    """Window holds exactly prev_frames frames; eviction removes the first sizes_.front() points."""
    om = oracle.LocalMapManager(3)
    frames = [np.full((n, 4), i, np.float32) for i, n in enumerate([5, 7, 0, 4, 6, 2])]
    for i, f in enumerate(frames):
        om.add(f)
        w, nf = om.get()
        keep = frames[max(0, i - 2):i + 1]
        assert nf == len(keep) and np.array_equal(w, np.concatenate(keep))


def test_odometer_first_frame_prediction_and_window():
    scans, gt = get_sequence("hdl64_small", 1000, 6)
    p = oracle.make_params(prev_frames=5)
    edges = _edges_of(p, scans)
    odo = oracle.Odometer(p)
    pose0, d0 = odo.process(edges[0])
    assert np.array_equal(pose0, np.eye(4)) and d0.n_matches[0] == 0     # first frame: identity, no solve
    w, nf = odo.window()
    assert nf == 1 and np.array_equal(w, edges[0])                        # raw edges seed the window
    poses = [pose0]
    for f in range(1, 6):
        o_odom, o_prev = odo.get_pose()
        pose, d = odo.process(edges[f])
        pred = o_odom @ (np.linalg.inv(o_prev) @ o_odom)                  # constant velocity (:148-150)
        assert np.allclose(np.array(d.pred_pose).reshape(4, 4), pred, atol=1e-12)
        assert d.n_map[0] == sum(len(e) for e in edges[max(0, f - 5):f])
        poses.append(pose)
    gold = np.load(os.path.join(GOLD, "trajectory_hdl64_small.npz"))["poses"]
    assert np.array_equal(np.stack(poses), gold)
    rel = np.linalg.inv(gt[0]) @ gt[5]
    dt, dr = pose_err(poses[5], rel)
    assert dt < 0.25 and dr < 0.02                                        # it actually tracks the motion


def test_voxelgrid_matches_numpy():
    rng = np.random.default_rng(3)
    pts = (rng.normal(size=(3000, 4)) * [3, 3, 1, 1]).astype(np.float32)
    o = oracle.voxelgrid(pts, 0.4)
    n = np_ref.voxelgrid(pts, 0.4)
    assert o.shape == n.shape and np.array_equal(o.view(np.uint32), n.view(np.uint32))
    assert len(oracle.voxelgrid(np.zeros((0, 4), np.float32), 0.4)) == 0


def test_map_keys_update_and_local_extraction():
    """Cell key int(floor(p*inv)*size + size/2) (src/map.cc:103-105); every touched cell is
    re-voxelised after each insert (:124-128); getLocalMap truncates the pose to int and its
    z-column loop never hits an existing cell for the shipped configs (:144-186)."""
    rng = np.random.default_rng(4)
    for (xy, z, cxy, cz) in ((40.0, 50.0, 2, 1), (30.0, 35.0, 3, 2), (20.0, 25.0, 2, 1)):
        om = oracle.Map(xy, z, 0.4)
        T = np.eye(4)
        allpts = []
        for f in range(6):
            pts = (rng.normal(size=(800, 4)) * [40, 40, 3, 1]).astype(np.float32)
            T[:3, 3] = [f * 5.0, -f * 2.0, 0.1 * f]
            om.update(pts, T)
            allpts.append(np_ref.transform(pts, T))
        keys, counts = om.cells()
        world = np.concatenate(allpts)
        exp_keys = {np_ref.map_cell_key(p, xy, z) for p in world[:, :3].astype(np.float64)}
        assert {tuple(k) for k in keys} == exp_keys
        assert counts.sum() == len(om.get_map()) and (counts > 0).all()
        # local map = the (2c+1)^2 square at the pose's z layer, concatenated i-outer j-inner
        T[:3, 3] = [7.9, -3.2, 0.4]
        loc = om.get_local_map(T, cxy, cz)
        kx, ky, kz = np_ref.map_cell_key([7.0, -3.0, 0.0], xy, z)    # translation truncated to int first
        full = om.get_map()
        offs = np.concatenate([[0], np.cumsum(counts)])
        exp = []
        for i in range(-cxy, cxy + 1):
            for j in range(-cxy, cxy + 1):
                for c, k in enumerate(keys):
                    if tuple(k) == (kx + int(i * xy), ky + int(j * xy), kz):
                        exp.append(full[offs[c]:offs[c + 1]])
        exp = np.concatenate(exp) if exp else np.zeros((0, 4), np.float32)
        assert np.array_equal(loc, exp)


def test_map_revoxelisation_is_unweighted_reaverage():
    """Old centroids re-enter the next VoxelGrid pass as single points (App. A.3)."""
    om = oracle.Map(40.0, 50.0, 0.4)
    a = np.array([[1.00, 1.0, 1.0, 0.0], [1.10, 1.0, 1.0, 0.0]], np.float32)
    om.update(a, np.eye(4))
    assert np.allclose(om.get_map()[0, 0], 1.05)
    om.update(np.array([[1.15, 1.0, 1.0, 0.0]], np.float32), np.eye(4))
    assert np.allclose(om.get_map()[0, 0], (1.05 + 1.15) / 2)     # not (1.0+1.1+1.15)/3
