"""ctypes front-end of oracle/_ref/libliodom_ref.so — TEST INFRASTRUCTURE ONLY.

The library is the REFERENCE'S OWN source (emiliofidalgo/liodom: src/feature_extractor.cc,
laser_odometry.cc, map.cc, params.cc, shared_data.cc, stats.cc) compiled unmodified from
/root/reference against the API shim in oracle/refshim/ (recipe: `make -C oracle ref`).  It exists to
pin oracle/liodom_oracle.cc; nothing under liodom_b200/ may import this module.

/root/reference is only present in the build container.  On the GPU box the prebuilt .so travels with
the repo; `available()` says whether it can be used.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libliodom_ref.so")
REFERENCE_ROOT = os.environ.get("LIODOM_REFERENCE_ROOT", "/root/reference")
_vp = ctypes.c_void_p
_lib = None


_DROPIN_NODE = os.path.join(_HERE, "_ref", "libliodom_dropin_node.so")
_DROPIN_MAPPING = os.path.join(_HERE, "_ref", "libliodom_dropin_mapping.so")


def build_dropin(force=False):
    """The reference's node mains (src/liodom_node.cc, src/liodom_mapping_node.cc), unmodified, over THIS repo's facade
    (needs liodom_b200/libliodom_b200.so to link).  Returns the two paths, or None when neither sources nor prebuilt
    libraries exist."""
    have = os.path.exists(_DROPIN_NODE) and os.path.exists(_DROPIN_MAPPING)
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        return (_DROPIN_NODE, _DROPIN_MAPPING) if have else None
    args = ["make", "-C", _HERE, "dropin", "CXX=g++", "REF=" + REFERENCE_ROOT]
    if force:
        args.insert(1, "-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return _DROPIN_NODE, _DROPIN_MAPPING


def dropin_available():
    return (os.path.exists(_DROPIN_NODE) and os.path.exists(_DROPIN_MAPPING)) or os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def _kv(params):
    out = []
    for k, v in params.items():
        if isinstance(v, bool):
            v = "true" if v else "false"
        out.append("%s=%s" % (k, v))
    return ";".join(out).encode()


def dropin_node_run(scans, width=0, height=0, dt=0.1, **params):
    """src/liodom_node.cc's main() (unmodified) over this repo's facade + CUDA library.
    -> (odom [n,13]: orientation x,y,z,w, position, twist linear, twist angular; edge counts [n]; frames produced)."""
    paths = build_dropin()
    L = ctypes.CDLL(paths[0])
    L.dropin_node_run.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_char_p, _vp, _vp]
    L.dropin_set_log_level(2)
    npts = np.array([len(s) for s in scans], np.int32)
    pts = np.ascontiguousarray(np.concatenate(scans), dtype=np.float32)
    odom = np.zeros((len(scans), 13))
    ne = np.zeros(len(scans), np.int32)
    n = L.dropin_node_run(_p(pts), _p(npts), len(scans), pts.shape[1], width, height, dt, _kv(params), _p(odom), _p(ne))
    return odom, ne, n


def dropin_mapping_run(clouds, poses, **params):
    """src/liodom_mapping_node.cc's main() (unmodified) over this repo's Map facade.
    -> (map_local sizes [n], last map_local [m,4], last map [k,4])."""
    paths = build_dropin()
    L = ctypes.CDLL(paths[1])
    L.dropin_mapping_run.argtypes = [_vp, _vp, ctypes.c_int, _vp, ctypes.c_char_p, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int]
    L.dropin_set_log_level(3)
    npts = np.array([len(c) for c in clouds], np.int32)
    pts = np.ascontiguousarray(np.concatenate(clouds), dtype=np.float32)
    P = np.ascontiguousarray(np.stack(poses).reshape(-1, 16), dtype=np.float64)
    sizes = np.zeros(len(clouds), np.int32)
    cap = int(npts.sum()) + 16
    loc = np.zeros((cap, 4), np.float32)
    mp = np.zeros((cap, 4), np.float32)
    k = L.dropin_mapping_run(_p(pts), _p(npts), len(clouds), _p(P), _kv(params), _p(sizes), _p(loc), cap, _p(mp), cap)
    return sizes, loc[:sizes[-1]].copy(), (mp[:k].copy() if k >= 0 else None)


def build(force=False):
    """Compile the reference's sources when they are present; otherwise keep a prebuilt library."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        return _SO if os.path.exists(_SO) else None
    args = ["make", "-C", _HERE, "ref", "CXX=g++", "REF=" + REFERENCE_ROOT]
    if force:
        args.insert(1, "-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return _SO


def available():
    return os.path.exists(_SO) or os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            raise RuntimeError("oracle/_ref is not built and %s is absent" % REFERENCE_ROOT)
        L = ctypes.CDLL(_SO)
        for f in ("ref_fext_create", "ref_lmap_create", "ref_odom_create", "ref_map_create"):
            getattr(L, f).restype = _vp
        L.ref_map_create.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double]
        L.ref_map_entropy.restype = ctypes.c_double
        L.ref_map_entropy.argtypes = [_vp]
        L.ref_warning_count.restype = ctypes.c_long
        L.ref_is_valid_point.argtypes = [_vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, _vp]
        L.ref_factor.argtypes = [_vp, _vp, _vp, ctypes.c_double, ctypes.c_double, _vp, _vp, _vp, _vp, _vp, _vp]
        L.ref_factor_residual.argtypes = [_vp, _vp, _vp, ctypes.c_double, ctypes.c_double, _vp, _vp, _vp]
        L.ref_odom_process.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_double, _vp]
        L.ref_run_sequence.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, _vp, _vp]
        L.ref_freeze_clock.argtypes = [ctypes.c_int, ctypes.c_double]
        L.ref_stats_add_times_ms.argtypes = [ctypes.c_double, ctypes.c_double]
        L.ref_stats_add_nfeats.argtypes = [ctypes.c_long]
        for f in ("ref_fext_destroy", "ref_lmap_destroy", "ref_odom_destroy", "ref_map_destroy", "ref_lmap_get", "ref_map_get",
                  "ref_odom_get_window"):
            getattr(L, f).argtypes = [_vp] if f.endswith("destroy") else [_vp, _vp]
        for f in ("ref_lmap_size", "ref_lmap_frames", "ref_map_size", "ref_map_num_cells", "ref_odom_window_size", "ref_odom_window_frames"):
            getattr(L, f).argtypes = [_vp]
        L.ref_lmap_add.argtypes = [_vp, _vp, ctypes.c_int]
        L.ref_lmap_set_max_frames.argtypes = [_vp, ctypes.c_int]
        L.ref_map_update.argtypes = [_vp, _vp, ctypes.c_int, _vp]
        L.ref_map_cells.argtypes = [_vp, _vp, _vp]
        L.ref_map_get_local.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int]
        L.ref_split.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp]
        L.ref_extract.argtypes = [_vp, _vp, _vp, _vp, ctypes.c_int]
        L.ref_fext_process.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int]
        L.ref_odom_get_pose.argtypes = [_vp, _vp, _vp]
        L.ref_odom_set_pose.argtypes = [_vp, _vp, _vp]
        L.ref_odom_set_window.argtypes = [_vp, _vp, _vp, ctypes.c_int]
        L.ref_set_received_map.argtypes = [_vp, ctypes.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _f4(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


def set_params(**kw):
    """ROS private params -> liodom::Params::readParams (src/params.cc:37-110).  Unset names take the
    reference's defaults.  A static identity transform laser -> base_link is registered so that
    getBaseToLaserTf (src/laser_odometry.cc:368-393) succeeds, as the launch files' static TF does."""
    L = lib()
    L.ref_clear_params()
    for k, v in kw.items():
        if isinstance(v, bool):
            v = "true" if v else "false"
        L.ref_set_param(k.encode(), str(v).encode())
    L.ref_read_params()
    L.ref_clear_static_tf()
    set_static_tf("velo_link", "base_link", [0, 0, 0], [0, 0, 0, 1])


def get_params():
    out = np.empty(13)
    lib().ref_get_params(_p(out))
    names = ("min_range", "max_range", "lidar_type", "scan_lines", "scan_regions", "edges_per_region", "min_points_per_scan",
             "local_map_size", "save_results", "use_imu", "filter_local_map", "mapping", "publish_tf")
    return dict(zip(names, out))


def set_static_tf(target, source, xyz, q_xyzw):
    lib().ref_set_static_tf(target.encode(), source.encode(), _p(np.asarray(xyz, np.float64)), _p(np.asarray(q_xyzw, np.float64)))


class FeatureExtractor:
    """liodom::FeatureExtractor of the reference (params must be set first)."""

    def __init__(self):
        self.h = _vp(lib().ref_fext_create())
        self.L = int(get_params()["scan_lines"])
        p = get_params()
        self.cap = int(p["scan_lines"] * p["scan_regions"] * (p["edges_per_region"] + 1))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_fext_destroy(self.h)
            self.h = None

    def is_valid_point(self, x, y, z):
        d = ctypes.c_double()
        ok = lib().ref_is_valid_point(self.h, x, y, z, ctypes.byref(d))
        return bool(ok), d.value

    def split(self, pts, width=0, height=0):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        n, stride = pts.shape
        rings = np.empty((n, 4), np.float32)
        off = np.empty(self.L + 1, np.int32)
        m = lib().ref_split(self.h, _p(pts), n, stride, width, height, _p(rings), _p(off))
        return dict(rings=rings[:m].copy(), offsets=off)

    def extract(self, rings, offsets):
        rings = _f4(rings)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        edges = np.empty((self.cap, 4), np.float32)
        e = lib().ref_extract(self.h, _p(rings), _p(offsets), _p(edges), self.cap)
        assert 0 <= e <= self.cap, e
        return edges[:e].copy()

    def process(self, pts, width=0, height=0):
        """The worker functor: pushPointCloud -> operator() -> popFeatures."""
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        edges = np.empty((self.cap, 4), np.float32)
        e = lib().ref_fext_process(self.h, _p(pts), pts.shape[0], pts.shape[1], width, height, _p(edges), self.cap)
        assert 0 <= e <= self.cap, e
        return edges[:e].copy()


class LocalMapManager:
    def __init__(self, max_frames):
        self.h = _vp(lib().ref_lmap_create(int(max_frames)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_lmap_destroy(self.h)
            self.h = None

    def add(self, pts):
        pts = _f4(pts)
        lib().ref_lmap_add(self.h, _p(pts), len(pts))

    def get(self):
        n = lib().ref_lmap_size(self.h)
        w = np.empty((n, 4), np.float32)
        lib().ref_lmap_get(self.h, _p(w))
        return w, lib().ref_lmap_frames(self.h)

    def set_max_frames(self, n):
        lib().ref_lmap_set_max_frames(self.h, int(n))


def factor(c, a, b, q, t, min_range=3.0, max_range=75.0):
    """Point2LineFactor through ceres::AutoDiffCostFunction -> (r[3], Jq[3,4], Jt[3,3], Jlocal[3,6])."""
    c, a, b, q, t = (np.ascontiguousarray(v, dtype=np.float64) for v in (c, a, b, q, t))
    r, Jq, Jt, Jl = np.empty(3), np.empty((3, 4)), np.empty((3, 3)), np.empty((3, 6))
    lib().ref_factor(_p(c), _p(a), _p(b), min_range, max_range, _p(q), _p(t), _p(r), _p(Jq), _p(Jt), _p(Jl))
    return r, Jq, Jt, Jl


def factor_residual(c, a, b, q, t, min_range=3.0, max_range=75.0):
    c, a, b, q, t = (np.ascontiguousarray(v, dtype=np.float64) for v in (c, a, b, q, t))
    r = np.empty(3)
    lib().ref_factor_residual(_p(c), _p(a), _p(b), min_range, max_range, _p(q), _p(t), _p(r))
    return r


class Odometer:
    """liodom::LaserOdometer of the reference, one popFeatures() iteration per process() call."""

    def __init__(self):
        self.h = _vp(lib().ref_odom_create())
        self.stamp = 1000.0

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_odom_destroy(self.h)
            self.h = None

    def process(self, edges, dt=0.1):
        edges = _f4(edges)
        pose = np.empty(16)
        rc = lib().ref_odom_process(self.h, _p(edges), len(edges), self.stamp, _p(pose))
        self.stamp += dt
        if rc != 0:
            raise RuntimeError("LaserOdometer returned without publishing (TF lookup failed?)")
        return pose.reshape(4, 4)

    def get_pose(self):
        o, q = np.empty(16), np.empty(16)
        lib().ref_odom_get_pose(self.h, _p(o), _p(q))
        return o.reshape(4, 4), q.reshape(4, 4)

    def set_pose(self, odom, prev_odom):
        o = np.ascontiguousarray(odom, dtype=np.float64).reshape(16)
        q = np.ascontiguousarray(prev_odom, dtype=np.float64).reshape(16)
        lib().ref_odom_set_pose(self.h, _p(o), _p(q))

    def window(self):
        n = lib().ref_odom_window_size(self.h)
        w = np.empty((n, 4), np.float32)
        lib().ref_odom_get_window(self.h, _p(w))
        return w, lib().ref_odom_window_frames(self.h)

    def set_window(self, pts, frame_sizes):
        pts = _f4(pts)
        fs = np.ascontiguousarray(frame_sizes, dtype=np.int32)
        assert fs.sum() == len(pts)
        lib().ref_odom_set_window(self.h, _p(pts), _p(fs), len(fs))

    @staticmethod
    def last_odom_msg():
        out = np.empty(13)
        seq = lib().ref_odom_last_msg(_p(out))
        return out, seq


def set_received_map(pts):
    pts = _f4(pts)
    lib().ref_set_received_map(_p(pts), len(pts))


def set_imu(q_xyzw):
    lib().ref_set_imu(_p(np.ascontiguousarray(q_xyzw, dtype=np.float64)))


def run_sequence(scans, width=0, height=0, dt=0.1):
    """The node pipeline (lidarClb -> FeatureExtractor worker -> LaserOdometer worker) over a list of scans.
    -> poses [n,4,4], edge counts [n]."""
    npts = np.array([len(s) for s in scans], np.int32)
    pts = np.ascontiguousarray(np.concatenate(scans), dtype=np.float32)
    poses = np.empty((len(scans), 16))
    ne = np.empty(len(scans), np.int32)
    n = lib().ref_run_sequence(_p(pts), _p(npts), len(scans), pts.shape[1], width, height, dt, _p(poses), _p(ne))
    assert n == len(scans), "pipeline stopped at frame %d" % n
    return poses.reshape(-1, 4, 4), ne


class Map:
    def __init__(self, xy_size=40.0, z_size=50.0, resolution=0.4):
        self.h = _vp(lib().ref_map_create(xy_size, z_size, resolution))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_map_destroy(self.h)
            self.h = None

    def update(self, pts, T):
        pts = _f4(pts)
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        lib().ref_map_update(self.h, _p(pts), len(pts), _p(T))

    def get_map(self):
        n = lib().ref_map_size(self.h)
        out = np.empty((n, 4), np.float32)
        lib().ref_map_get(self.h, _p(out))
        return out

    def cells(self):
        n = lib().ref_map_num_cells(self.h)
        keys = np.empty((n, 3), np.int32)
        counts = np.empty(n, np.int32)
        lib().ref_map_cells(self.h, _p(keys), _p(counts))
        return keys, counts

    def get_local_map(self, T, cells_xy=2, cells_z=1):
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        n = lib().ref_map_get_local(self.h, _p(T), cells_xy, cells_z, None, 0)
        out = np.empty((n, 4), np.float32)
        lib().ref_map_get_local(self.h, _p(T), cells_xy, cells_z, _p(out), n)
        return out

    def entropy(self):
        return lib().ref_map_entropy(self.h)


def stats_write(poses, nfeats, times_ms, directory):
    """Stats::addPose / addNumOfFeats / add*Time / writeResults (src/stats.cc) -> the five text files."""
    L = lib()
    L.ref_stats_clear()
    for T in poses:
        L.ref_stats_add_pose(_p(np.ascontiguousarray(T, dtype=np.float64).reshape(16)))
    for n in nfeats:
        L.ref_stats_add_nfeats(int(n))
    for a, b in times_ms:
        L.ref_stats_add_times_ms(float(a), float(b))
    d = directory if directory.endswith("/") else directory + "/"
    L.ref_stats_write(d.encode())
