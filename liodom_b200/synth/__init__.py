"""Synthetic LiDAR workloads: ctypes front-end of synth.cc.

Harness code, not part of the hot path: it only fabricates the inputs that tests and
bench.py feed to both the CUDA path and the oracle.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libliodom_synth.so")
_SRC = os.path.join(_HERE, "synth.cc")


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["g++", "-O3", "-fopenmp", "-shared", "-fPIC", "-std=c++17", "-o", _SO, _SRC])
    return _SO


class _Sensor(ctypes.Structure):
    _fields_ = [("model", ctypes.c_int), ("beams", ctypes.c_int), ("az_steps", ctypes.c_int),
                ("order", ctypes.c_int), ("noise_sigma", ctypes.c_double), ("max_ray", ctypes.c_double)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.synth_scan.restype = ctypes.c_int
        _lib.synth_scan.argtypes = [ctypes.POINTER(_Sensor), ctypes.c_uint64, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_double), ctypes.c_void_p]
        _lib.synth_pose.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    return _lib


# name -> (model, beams, az_steps, order)
SENSORS = {
    "hdl64": (0, 64, 1875, 0),          # C1/C3/C5: ring-major, ~120k returns
    "hdl64_firing": (0, 64, 1875, 1),   # same rays in firing (azimuth-major) order
    "os1_128": (1, 128, 2048, 2),       # C2: organised 128 x 2048, holes = (0,0,0)
    "hdl64_1m": (0, 64, 15625, 0),      # C5: 1M-point scan
    "hdl64_small": (0, 64, 400, 0),     # reduced azimuth resolution for quick CPU tests
}


def gt_pose(seed, frame, traj=0):
    """Ground-truth world_from_sensor pose, 4x4 float64."""
    lib = _load()
    T = (ctypes.c_double * 16)()
    lib.synth_pose(ctypes.c_uint64(seed), traj, frame, T)
    return np.array(T, dtype=np.float64).reshape(4, 4)


def scan(sensor, seed, frame, traj=0, noise_sigma=0.02, max_ray=85.0, pose=None):
    """One scan as float32 [n,4] (x,y,z,intensity) in the sensor frame.

    Organised sensors return all beams*az_steps slots (row-major, missing = zeros).
    """
    lib = _load()
    model, beams, az, order = SENSORS[sensor]
    s = _Sensor(model, beams, az, order, noise_sigma, max_ray)
    T = gt_pose(seed, frame, traj) if pose is None else np.ascontiguousarray(pose, dtype=np.float64)
    buf = np.zeros((beams * az, 4), dtype=np.float32)
    n = lib.synth_scan(ctypes.byref(s), ctypes.c_uint64(seed), frame,
                       T.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), buf.ctypes.data_as(ctypes.c_void_p))
    return buf[:n]


def sensor_shape(sensor):
    """(width, height) of the organised layout (height = beams)."""
    _, beams, az, _ = SENSORS[sensor]
    return az, beams


def sequence(sensor, seed, nframes, traj=0, start=0, **kw):
    """List of scans and the matching ground-truth poses [nframes,4,4]."""
    scans = [scan(sensor, seed, start + f, traj, **kw) for f in range(nframes)]
    poses = np.stack([gt_pose(seed, start + f, traj) for f in range(nframes)])
    return scans, poses
