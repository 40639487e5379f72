#!/usr/bin/env python
"""Generates the committed golden fixtures from the CPU oracle on seeded synthetic inputs.

The reference (emiliofidalgo/liodom) ships no tests, fixtures or golden vectors and cannot be
built in this image (ROS/PCL/FLANN/Ceres/Eigen absent), so these vectors are outputs of the
oracle restatement at the time of generation — they pin the oracle against regressions and let
the GPU tests check against committed numbers; they are not outputs of the reference binary.

    PYTHONPATH=. python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from liodom_b200 import synth  # noqa: E402


def main():
    seed = 1000
    # 1. extraction on one reduced-resolution HDL-64 scan
    scans, gt = synth.sequence("hdl64_small", seed, 6)
    p = oracle.make_params()
    sp = oracle.split(p, scans[0])
    ex = oracle.extract(p, sp["rings"], sp["offsets"], want_keys=True)
    np.savez_compressed(os.path.join(HERE, "extract_hdl64_small.npz"), seed=seed, scan=scans[0], offsets=sp["offsets"],
                        edge_ring=ex["ring"], edge_idx=ex["idx"], edges=ex["edges"], keys_nansum=np.nansum(ex["keys"]))
    # 2. association + one solve, teacher-forced: window = GT-posed edges of frames 0..3
    edges = []
    for s in scans:
        q = oracle.split(p, s)
        edges.append(oracle.extract(p, q["rings"], q["offsets"])["edges"])
    rel = [np.linalg.inv(gt[0]) @ g for g in gt]
    window = np.concatenate([oracle.transform(edges[k], rel[k]) for k in range(4)])
    T = rel[4].copy()
    T[:3, 3] += [0.08, -0.05, 0.01]
    a = oracle.associate(edges[4], T, window, knn_method=1)
    sel = (a["gate"] & 2) == 2
    cab = np.concatenate([edges[4][sel][:, :3], window[a["knn_idx"][sel][:, 0]][:, :3], window[a["knn_idx"][sel][:, 1]][:, :3]], 1).astype(np.float64)
    from scipy.spatial.transform import Rotation
    q0 = Rotation.from_matrix(T[:3, :3]).as_quat()
    q0 = -q0 if q0[3] < 0 else q0
    q1, t1, summ = oracle.solve(cab, q0, T[:3, 3])
    np.savez_compressed(os.path.join(HERE, "register_hdl64_small.npz"), seed=seed, frame_sizes=[len(e) for e in edges[:4]],
                        window=window, edges=edges[4], pose=T, knn_idx=a["knn_idx"], knn_d2=a["knn_d2"], gate=a["gate"],
                        tie=a["tie"], eig=a["eig"], cab=cab, q0=q0, t0=T[:3, 3], q1=q1, t1=t1,
                        summary=[summ.iterations, summ.successful_steps, summ.termination, summ.num_residual_blocks],
                        costs=[summ.initial_cost, summ.final_cost])
    # 3. a free-running 6-frame trajectory (poses only)
    poses, _, _ = oracle.run_sequence(oracle.make_params(prev_frames=5), scans)
    np.savez_compressed(os.path.join(HERE, "trajectory_hdl64_small.npz"), seed=seed, poses=poses)
    # 4. the steps either side of the path: window filter, IMU override, odometry message, PointCloud2 decode
    K = 4
    pf = oracle.make_params(prev_frames=K, filter_local_map=1)
    poses_f, _, _ = oracle.run_sequence(pf, scans)
    window4 = np.concatenate([oracle.transform(edges[k], rel[k]) for k in range(K)])
    filt = oracle.voxelgrid(window4, 0.4)
    l2b = np.eye(4)
    l2b[:3, :3] = Rotation.from_euler("ZYX", [0.3, -0.015, 0.01]).as_matrix()
    l2b[:3, 3] = [0.2, 0.1, -1.5]
    imu_q = Rotation.from_euler("ZYX", [0.7, 0.02, -0.03]).as_quat()
    over = oracle.imu_override(rel[3], imu_q, l2b)
    odo = oracle.publish_odom(rel[3], rel[2], l2b, 0.1)
    rng = np.random.default_rng(seed)
    blob = rng.integers(0, 256, 22 * 64, dtype=np.uint8)
    dec = oracle.decode_cloud2(blob, 16, 4, 22, 16 * 22, 0, 4, 8, 12)
    np.savez_compressed(os.path.join(HERE, "extras_hdl64_small.npz"), seed=seed, poses_filtered=poses_f, filtered_window=filt,
                        l2b=l2b, imu_q=imu_q, pose3=rel[3], pose2=rel[2], imu_override=over, odometry=odo, blob=blob, decoded=dec)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
