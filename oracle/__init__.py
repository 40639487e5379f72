"""ctypes front-end of the CPU oracle (liodom_oracle.cc).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never from liodom_b200/.
See liodom_oracle.h for what is restated from where and the "parity unpinned" note.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libliodom_oracle.so")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("liodom_oracle.cc", "liodom_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "CXX=g++"], stdout=subprocess.DEVNULL)
    return _SO


class Params(ctypes.Structure):
    _fields_ = [("min_range", ctypes.c_double), ("max_range", ctypes.c_double), ("lidar_type", ctypes.c_int),
                ("scan_lines", ctypes.c_int), ("scan_regions", ctypes.c_int), ("edges_per_region", ctypes.c_int),
                ("prev_frames", ctypes.c_int), ("filter_local_map", ctypes.c_int), ("mapping", ctypes.c_int),
                ("omp_threads", ctypes.c_int)]


class SolveSummary(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_int), ("successful_steps", ctypes.c_int), ("termination", ctypes.c_int),
                ("initial_cost", ctypes.c_double), ("final_cost", ctypes.c_double),
                ("num_residual_blocks", ctypes.c_int), ("cost_evals", ctypes.c_int), ("jac_evals", ctypes.c_int)]


class FrameDiag(ctypes.Structure):
    _fields_ = [("n_edges", ctypes.c_int), ("n_map", ctypes.c_int * 2), ("n_matches", ctypes.c_int * 2),
                ("solve", SolveSummary * 2), ("pred_pose", ctypes.c_double * 16), ("times_us", ctypes.c_double * 4)]


_lib = None
_vp = ctypes.c_void_p


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.orc_split.restype = ctypes.c_int
        L.orc_extract.restype = ctypes.c_int
        L.orc_extract_scan.restype = ctypes.c_int
        L.orc_voxelgrid.restype = ctypes.c_int
        L.orc_voxelgrid.argtypes = [_vp, ctypes.c_int, ctypes.c_float, _vp]
        for f in ("orc_odom_create", "orc_lmap_create", "orc_map_create"):
            getattr(L, f).restype = _vp
        L.orc_map_create.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double]
        L.orc_solve.argtypes = [_vp, ctypes.c_int, ctypes.c_double, ctypes.c_double, _vp, _vp, ctypes.c_int, _vp]
        L.orc_factor.argtypes = [_vp, _vp, _vp, ctypes.c_double, ctypes.c_double, _vp, _vp, _vp, _vp]
        L.orc_run_sequence.restype = ctypes.c_long
        for f in ("orc_odom_window_size", "orc_odom_window_frames", "orc_lmap_size", "orc_lmap_frames",
                  "orc_map_size", "orc_map_num_cells", "orc_map_get_local"):
            getattr(L, f).restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _f4(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


def make_params(**kw):
    p = Params()
    lib().orc_default_params(ctypes.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def split(p, pts, width=0, height=0):
    """-> dict(ring_of_point, rings [nv,4], offsets [L+1], src_index [nv], status)."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    n, stride = pts.shape
    ring = np.empty(n, np.int32)
    rings = np.empty((n, 4), np.float32)
    off = np.empty(p.scan_lines + 1, np.int32)
    src = np.empty(n, np.int32)
    nv = lib().orc_split(ctypes.byref(p), _p(pts), n, stride, width, height, _p(ring), _p(rings), _p(off), _p(src))
    m = off[-1]
    return dict(ring_of_point=ring, rings=rings[:m].copy(), offsets=off, src_index=src[:m].copy(), status=nv)


def extract(p, rings, offsets, sort_mode=0, want_keys=False):
    """-> dict(edges [E,4], ring [E], idx [E], keys [nv] or None)."""
    rings = _f4(rings)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    cap = p.scan_lines * p.scan_regions * (p.edges_per_region + 1)
    edges = np.empty((cap, 4), np.float32)
    er = np.empty(cap, np.int32)
    ei = np.empty(cap, np.int32)
    keys = np.empty(len(rings), np.float64) if want_keys else None
    e = lib().orc_extract(ctypes.byref(p), _p(rings), _p(offsets), _p(edges), _p(er), _p(ei), _p(keys), sort_mode, cap)
    assert e <= cap
    return dict(edges=edges[:e].copy(), ring=er[:e].copy(), idx=ei[:e].copy(), keys=keys)


def extract_scan(p, pts, width=0, height=0):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    cap = p.scan_lines * p.scan_regions * (p.edges_per_region + 1)
    edges = np.empty((cap, 4), np.float32)
    t = np.zeros(2)
    e = lib().orc_extract_scan(ctypes.byref(p), _p(pts), pts.shape[0], pts.shape[1], width, height, _p(edges), cap, _p(t))
    return edges[:e].copy(), t


def transform(pts, T):
    pts = _f4(pts)
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
    out = np.empty_like(pts)
    lib().orc_transform(_p(pts), len(pts), _p(T), _p(out))
    return out


def knn5(map_pts, q, method=0):
    map_pts = _f4(map_pts)
    q = _f4(q)
    E = len(q)
    idx = np.empty((E, 5), np.int32)
    d2 = np.empty((E, 5), np.float32)
    tie = np.empty(E, np.uint8)
    lib().orc_knn5(_p(map_pts), len(map_pts), _p(q), E, method, _p(idx), _p(d2), _p(tie))
    return idx, d2, tie


def associate(edges, T, map_pts, knn_method=0):
    edges = _f4(edges)
    map_pts = _f4(map_pts)
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
    E = len(edges)
    out = dict(knn_idx=np.empty((E, 5), np.int32), knn_d2=np.empty((E, 5), np.float32), gate=np.empty(E, np.uint8),
               eig=np.empty((E, 3), np.float64), q_world=np.empty((E, 4), np.float32), tie=np.empty(E, np.uint8))
    lib().orc_associate(_p(edges), E, _p(T), _p(map_pts), len(map_pts), knn_method, _p(out["knn_idx"]),
                        _p(out["knn_d2"]), _p(out["gate"]), _p(out["eig"]), _p(out["q_world"]), _p(out["tie"]))
    return out


def factor(c, a, b, q, t, min_range=3.0, max_range=75.0):
    c, a, b, q, t = (np.ascontiguousarray(v, dtype=np.float64) for v in (c, a, b, q, t))
    r = np.empty(3)
    J = np.empty((3, 6))
    lib().orc_factor(_p(c), _p(a), _p(b), min_range, max_range, _p(q), _p(t), _p(r), _p(J))
    return r, J


def solve(cab, q, t, min_range=3.0, max_range=75.0, linear_solver=0):
    cab = np.ascontiguousarray(cab, dtype=np.float64).reshape(-1, 9)
    q = np.array(q, dtype=np.float64)
    t = np.array(t, dtype=np.float64)
    s = SolveSummary()
    lib().orc_solve(_p(cab), len(cab), min_range, max_range, _p(q), _p(t), linear_solver, ctypes.byref(s))
    return q, t, s


class Odometer:
    """LaserOdometer restated (state machine around orc_odom_*)."""

    def __init__(self, p):
        self.p = p
        self.h = _vp(lib().orc_odom_create(ctypes.byref(p)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_odom_destroy(self.h)
            self.h = None

    def process(self, edges):
        edges = _f4(edges)
        pose = np.empty(16)
        d = FrameDiag()
        lib().orc_odom_process(self.h, _p(edges), len(edges), _p(pose), ctypes.byref(d))
        return pose.reshape(4, 4), d

    def set_pose(self, odom, prev_odom):
        o = np.ascontiguousarray(odom, dtype=np.float64).reshape(16)
        q = np.ascontiguousarray(prev_odom, dtype=np.float64).reshape(16)
        lib().orc_odom_set_pose(self.h, _p(o), _p(q))

    def get_pose(self):
        o = np.empty(16)
        q = np.empty(16)
        lib().orc_odom_get_pose(self.h, _p(o), _p(q))
        return o.reshape(4, 4), q.reshape(4, 4)

    def window(self):
        n = lib().orc_odom_window_size(self.h)
        w = np.empty((n, 4), np.float32)
        lib().orc_odom_get_window(self.h, _p(w))
        return w, lib().orc_odom_window_frames(self.h)

    def set_window(self, pts, frame_sizes):
        pts = _f4(pts)
        fs = np.ascontiguousarray(frame_sizes, dtype=np.int32)
        assert fs.sum() == len(pts)
        lib().orc_odom_set_window(self.h, _p(pts), _p(fs), len(fs))

    def set_imu(self, use_imu, q_xyzw=None, laser_to_base=None):
        q = None if q_xyzw is None else np.ascontiguousarray(q_xyzw, dtype=np.float64)
        T = None if laser_to_base is None else np.ascontiguousarray(laser_to_base, dtype=np.float64)
        lib().orc_odom_set_imu(self.h, int(use_imu), _p(q), _p(T))

    def set_received_map(self, pts):
        pts = _f4(pts)
        lib().orc_odom_set_received_map(self.h, _p(pts), len(pts))


class LocalMapManager:
    def __init__(self, max_frames):
        self.h = _vp(lib().orc_lmap_create(int(max_frames)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_lmap_destroy(self.h)
            self.h = None

    def add(self, pts):
        pts = _f4(pts)
        lib().orc_lmap_add(self.h, _p(pts), len(pts))

    def get(self):
        n = lib().orc_lmap_size(self.h)
        w = np.empty((n, 4), np.float32)
        lib().orc_lmap_get(self.h, _p(w))
        return w, lib().orc_lmap_frames(self.h)

    def set_max_frames(self, n):
        lib().orc_lmap_set_max_frames(self.h, int(n))


def decode_cloud2(data, width, height, point_step, row_step, off_x, off_y, off_z, off_i):
    """pcl::fromROSMsg<PointXYZI> of a raw sensor_msgs/PointCloud2 blob -> [n,4] float32."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    out = np.empty((width * height, 4), np.float32)
    lib().orc_decode_cloud2(_p(data), width, height, point_step, row_step, off_x, off_y, off_z, off_i, _p(out))
    return out


def tf_rpy(q_xyzw):
    """-> (rpy [3], quaternion rebuilt from them) through tf::Matrix3x3 getRPY / setRPY / getRotation."""
    q = np.ascontiguousarray(q_xyzw, dtype=np.float64)
    rpy = np.empty(3)
    back = np.empty(4)
    lib().orc_tf_rpy(_p(q), _p(rpy), _p(back))
    return rpy, back


def imu_override(odom, imu_q_xyzw, laser_to_base):
    out = np.empty((4, 4))
    lib().orc_imu_override(_p(np.ascontiguousarray(odom, dtype=np.float64)), _p(np.ascontiguousarray(imu_q_xyzw, dtype=np.float64)),
                           _p(np.ascontiguousarray(laser_to_base, dtype=np.float64)), _p(out))
    return out


def publish_odom(pose, prev_odom, laser_to_base, delta_time):
    """-> 13 doubles: orientation x,y,z,w, position, twist linear, twist angular (publishOdom)."""
    out = np.empty(13)
    lib().orc_publish_odom.argtypes = [_vp, _vp, _vp, ctypes.c_double, _vp]
    lib().orc_publish_odom(_p(np.ascontiguousarray(pose, dtype=np.float64)), _p(np.ascontiguousarray(prev_odom, dtype=np.float64)),
                           _p(np.ascontiguousarray(laser_to_base, dtype=np.float64)), float(delta_time), _p(out))
    return out


def voxelgrid(pts, leaf):
    pts = _f4(pts)
    out = np.empty_like(pts)
    n = lib().orc_voxelgrid(_p(pts), len(pts), leaf, _p(out))
    return out[:max(n, 0)].copy() if n >= 0 else pts.copy()


class Map:
    def __init__(self, xy_size=40.0, z_size=50.0, resolution=0.4):
        self.h = _vp(lib().orc_map_create(xy_size, z_size, resolution))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_map_destroy(self.h)
            self.h = None

    def update(self, pts, T):
        pts = _f4(pts)
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        lib().orc_map_update(self.h, _p(pts), len(pts), _p(T))

    def get_map(self):
        n = lib().orc_map_size(self.h)
        out = np.empty((n, 4), np.float32)
        lib().orc_map_get(self.h, _p(out))
        return out

    def cells(self):
        n = lib().orc_map_num_cells(self.h)
        keys = np.empty((n, 3), np.int32)
        counts = np.empty(n, np.int32)
        for i in range(n):
            c = ctypes.c_int32()
            lib().orc_map_cell_info(self.h, i, _p(keys[i:i + 1]), ctypes.byref(c))
            counts[i] = c.value
        return keys, counts

    def get_local_map(self, T, cells_xy=2, cells_z=1):
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        n = lib().orc_map_get_local(self.h, _p(T), cells_xy, cells_z, None, 0)
        out = np.empty((n, 4), np.float32)
        lib().orc_map_get_local(self.h, _p(T), cells_xy, cells_z, _p(out), n)
        return out


def run_sequence(p, scans, width=0, height=0):
    """CPU baseline: whole path over a list of scans. -> poses [n,4,4], stage_us[5], edges."""
    npts = np.array([len(s) for s in scans], np.int32)
    pts = np.ascontiguousarray(np.concatenate(scans), dtype=np.float32)
    poses = np.empty((len(scans), 16))
    st = np.zeros(5)
    e = lib().orc_run_sequence(ctypes.byref(p), _p(pts), _p(npts), len(scans), pts.shape[1], width, height,
                               _p(poses), _p(st))
    return poses.reshape(-1, 4, 4), st, e
