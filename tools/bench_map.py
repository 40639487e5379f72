#!/usr/bin/env python
"""Config C4 (BASELINE.json): global hash-grid map build over a synthetic closed loop — per frame
Map::updateMap + Map::getLocalMap (src/liodom_mapping_node.cc:45-90), GPU (C ABI) vs the oracle Map
on the host.  Edge clouds come from the GPU extractor, poses are the ground truth.

    python tools/bench_map.py [--frames 400] [--sensor hdl64_small]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (CPU baseline leg only)
from liodom_b200 import api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=400)
    ap.add_argument("--sensor", default="hdl64_small")
    ap.add_argument("--stride", type=int, default=10, help="take every stride-th frame of the 1 m/frame circuit")
    a = ap.parse_args()
    ctx = api.Context(max_points=131072)
    clouds, poses = [], []
    T0 = synth.gt_pose(1000, 0, traj=1)
    for k in range(a.frames):
        f = k * a.stride
        clouds.append(ctx.extract(synth.scan(a.sensor, 1000, f, traj=1)))
        poses.append(np.linalg.inv(T0) @ synth.gt_pose(1000, f, traj=1))
    ctx.close()
    out = {}
    for name, (xy, z, cxy, cz) in {"mapping.launch": (20.0, 25.0, 2, 1), "liodom.launch": (30.0, 35.0, 3, 2)}.items():
        gm = api.Map(xy, z, 0.4, max_points=1 << 22)
        gm.update(clouds[0], poses[0])          # warm-up (allocations, first kernels)
        gm.close()
        gm = api.Map(xy, z, 0.4, max_points=1 << 22)
        t0 = time.perf_counter()
        nloc = 0
        for c, T in zip(clouds, poses):
            gm.update(c, T)
            nloc += len(gm.get_local_map(T, cxy, cz))
        tg = time.perf_counter() - t0
        npts, ncells = gm.size()
        gfull = gm.get_map()
        om = oracle.Map(xy, z, 0.4)
        t0 = time.perf_counter()
        for c, T in zip(clouds, poses):
            om.update(c, T)
            om.get_local_map(T, cxy, cz)
        tc = time.perf_counter() - t0
        same = bool(np.array_equal(gfull.view(np.uint32), om.get_map().view(np.uint32)))
        out[name] = {"voxel_xy": xy, "voxel_z": z, "cells_xy": cxy, "cells_z": cz, "frames": a.frames,
                     "edges_per_frame": int(np.mean([len(c) for c in clouds])), "map_points": npts, "cells": ncells,
                     "local_map_points_per_frame": int(nloc / a.frames),
                     "gpu_ms_per_frame": round(tg / a.frames * 1e3, 3), "cpu_oracle_ms_per_frame": round(tc / a.frames * 1e3, 3),
                     "map_bitwise_equal_to_oracle": same}
        gm.close()
    print(json.dumps({"bench": "C4_map_build", "sensor": a.sensor, "results": out}))


if __name__ == "__main__":
    main()
