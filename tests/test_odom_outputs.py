"""Outputs and IMU hand-off on the step after / before the path:
publishOdom's pose-in-base_link and twist arithmetic (src/laser_odometry.cc:395-446) and the use_imu
roll/pitch override (src/laser_odometry.cc:152-183).  CPU: the oracle's tf restatement against
SciPy and the façade's arithmetic against the oracle; GPU: the on-device override against the oracle."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import oracle
from conftest import get_sequence, pose_err


def _T(rpy, t):
    T = np.eye(4)
    T[:3, :3] = Rotation.from_euler("ZYX", [rpy[2], rpy[1], rpy[0]]).as_matrix()   # Rz(yaw) Ry(pitch) Rx(roll), tf's convention
    T[:3, 3] = t
    return T


def test_tf_rpy_matches_scipy():
    rng = np.random.default_rng(0)
    for _ in range(200):
        q = rng.normal(size=4)
        rpy, back = oracle.tf_rpy(q)
        R = Rotation.from_quat(q / np.linalg.norm(q))
        ypr = R.as_euler("ZYX")
        assert np.allclose(rpy, [ypr[2], ypr[1], ypr[0]], atol=1e-12)
        qn = q / np.linalg.norm(q)
        assert min(np.abs(back - qn).max(), np.abs(back + qn).max()) < 1e-12
    # gimbal lock: pitch = +-pi/2 -> yaw forced to 0
    for sgn in (1.0, -1.0):
        q = Rotation.from_euler("ZYX", [0.3, sgn * np.pi / 2, 0.2]).as_quat()
        rpy, _ = oracle.tf_rpy(q)
        assert abs(abs(rpy[1]) - np.pi / 2) < 1e-7


def test_imu_override_keeps_yaw_and_position_takes_roll_pitch():
    rng = np.random.default_rng(1)
    l2b = _T([0.01, -0.02, 0.5], [0.3, -0.1, -1.2])
    for _ in range(50):
        odom = _T(rng.normal(size=3) * [0.1, 0.1, 1.0], rng.normal(size=3) * 10)
        imu = Rotation.from_euler("ZYX", [rng.normal(), rng.normal() * 0.1, rng.normal() * 0.1])
        out = oracle.imu_override(odom, imu.as_quat(), l2b)
        bl_in, bl_out = odom @ l2b, out @ l2b
        ypr_in = Rotation.from_matrix(bl_in[:3, :3]).as_euler("ZYX")
        ypr_out = Rotation.from_matrix(bl_out[:3, :3]).as_euler("ZYX")
        ypr_imu = imu.as_euler("ZYX")
        assert abs(ypr_out[0] - ypr_in[0]) < 1e-9                      # yaw kept
        assert np.allclose(ypr_out[1:], ypr_imu[1:], atol=1e-9)         # pitch, roll from the IMU
        assert np.allclose(bl_out[:3, 3], bl_in[:3, 3], atol=1e-9)      # base_link position unchanged
    # identity laser_to_base and an IMU equal to the pose's own roll/pitch: nothing changes
    odom = _T([0.02, -0.03, 0.7], [1, 2, 3])
    out = oracle.imu_override(odom, Rotation.from_matrix(odom[:3, :3]).as_quat(), np.eye(4))
    assert np.allclose(out, odom, atol=1e-12)


def test_publish_odom_arithmetic_facade_vs_oracle_vs_scipy():
    from liodom_b200 import host_api
    rng = np.random.default_rng(2)
    l2b = _T([0.0, 0.0, 0.1], [0.5, 0.0, -1.0])
    for _ in range(50):
        prev = _T(rng.normal(size=3) * [0.05, 0.05, 1.0], rng.normal(size=3) * 20)
        step = _T(rng.normal(size=3) * 0.02, rng.normal(size=3) * 0.5 + [1.0, 0, 0])
        pose = prev @ step
        dt = 0.1
        o = oracle.publish_odom(pose, prev, l2b, dt)
        f = host_api.make_odometry(pose, prev, l2b, 5.0 + dt, 5.0)
        assert np.allclose(f, o, rtol=0, atol=1e-9)
        bl = pose @ l2b
        qs = Rotation.from_matrix(bl[:3, :3]).as_quat()
        assert min(np.abs(o[:4] - qs).max(), np.abs(o[:4] + qs).max()) < 1e-12
        assert np.allclose(o[4:7], bl[:3, 3])
        delta = np.linalg.inv(prev @ l2b) @ bl
        assert np.allclose(o[7:10], delta[:3, 3] / dt)
        ypr = Rotation.from_matrix(delta[:3, :3]).as_euler("ZYX")
        assert np.allclose(o[10:13], np.array([ypr[2], ypr[1], ypr[0]]) / dt, atol=1e-9)


@pytest.mark.gpu
def test_gpu_use_imu_matches_oracle(cuda_lib):
    """use_imu=1 with a tilted laser->base transform: per-frame IMU orientations (ground-truth roll and
    pitch + noise) go to both sides; poses agree within 1e-4 m / 1e-5 rad teacher-forced."""
    from liodom_b200 import api
    scans, gt = get_sequence("hdl64_small", 1003, 8)
    op = oracle.make_params(prev_frames=5)
    l2b = _T([0.01, -0.015, 0.3], [0.2, 0.1, -1.5])
    rng = np.random.default_rng(7)
    ctx = api.Context(prev_frames=5, max_points=32768, use_imu=1)
    ctx.set_laser_to_base(l2b)
    odo = oracle.Odometer(op)
    sizes = []
    for f, s in enumerate(scans):
        sp = oracle.split(op, s)
        edges = oracle.extract(op, sp["rings"], sp["offsets"])["edges"]
        bl = (np.linalg.inv(gt[0]) @ gt[f]) @ l2b
        imu_q = (Rotation.from_matrix(bl[:3, :3]) * Rotation.from_euler("ZYX", rng.normal(size=3) * [0.5, 0.002, 0.002])).as_quat()
        odo.set_imu(1, imu_q, l2b)
        ctx.set_imu(imu_q)
        if f > 0:   # teacher forcing: the oracle's pre-frame state into the GPU lane
            w, nf = odo.window()
            ctx.lmap_clear()
            pos = 0
            for n in sizes:
                ctx.lmap_add(w[pos:pos + n])
                pos += n
            ctx.set_pose(*odo.get_pose())
        opose, od = odo.process(edges)
        gpose, gd = ctx.register(edges)
        if f > 0:
            pp = np.abs(np.array(gd.pred_pose).reshape(4, 4) - np.array(od.pred_pose).reshape(4, 4)).max()
            assert pp < 1e-12, pp     # the override runs on device libm: a few ulp, not bit-equal
        dt, dr = pose_err(gpose, opose)
        assert dt < 1e-4 and dr < 1e-5, (f, dt, dr)
        sizes = (sizes + [len(edges)])[-5:]
    # the override must actually have changed something relative to use_imu=0
    ctx.close()


def _extras():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "extras_hdl64_small.npz"))


def test_golden_extras_pin_the_oracle():
    """Committed vectors (tests/golden/make_golden.py) for the steps either side of the path: window filter,
    IMU override, odometry message, PointCloud2 decode."""
    g = _extras()
    scans, gt = get_sequence("hdl64_small", int(g["seed"]), 6)
    op = oracle.make_params()
    edges = []
    for s in scans:
        sp = oracle.split(op, s)
        edges.append(oracle.extract(op, sp["rings"], sp["offsets"])["edges"])
    rel = [np.linalg.inv(gt[0]) @ p for p in gt]
    window = np.concatenate([oracle.transform(edges[k], rel[k]) for k in range(4)])
    assert np.array_equal(oracle.voxelgrid(window, 0.4).view(np.uint32), g["filtered_window"].view(np.uint32))
    poses, _, _ = oracle.run_sequence(oracle.make_params(prev_frames=4, filter_local_map=1), scans)
    assert np.allclose(poses, g["poses_filtered"], rtol=0, atol=1e-12)
    assert np.allclose(oracle.imu_override(g["pose3"], g["imu_q"], g["l2b"]), g["imu_override"], rtol=0, atol=1e-14)
    assert np.allclose(oracle.publish_odom(g["pose3"], g["pose2"], g["l2b"], 0.1), g["odometry"], rtol=0, atol=1e-13)
    dec = oracle.decode_cloud2(g["blob"], 16, 4, 22, 16 * 22, 0, 4, 8, 12)
    assert np.array_equal(dec.view(np.uint32), g["decoded"].view(np.uint32))


@pytest.mark.gpu
def test_gpu_against_golden_extras(cuda_lib):
    """The CUDA path against the committed vectors: filtered-window trajectory and the decoded cloud's edges."""
    from liodom_b200 import api
    g = _extras()
    scans, _ = get_sequence("hdl64_small", int(g["seed"]), 6)
    ctx = api.Context(prev_frames=4, filter_local_map=1, max_points=32768)
    for f, s in enumerate(scans):
        ctx.scan_batch([s])
        p, _ = ctx.results()
        dt, dr = pose_err(p[0], g["poses_filtered"][f])
        assert dt < 1e-3 and dr < 1e-4, (f, dt, dr)
    ctx.close()
