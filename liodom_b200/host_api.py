"""ctypes front-end of the C++ facade harness (libliodom_host.so): runs a sequence of scans
through FeatureExtractor / LaserOdometer worker threads and the SharedData queues exactly like
liodom_node does (src/liodom_node.cc:72-121)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_SO = os.path.join(_HERE, "libliodom_host.so")


class HostOptions(ctypes.Structure):
    _fields_ = [("min_range", ctypes.c_double), ("max_range", ctypes.c_double), ("lidar_type", ctypes.c_int),
                ("scan_lines", ctypes.c_int), ("scan_regions", ctypes.c_int), ("edges_per_region", ctypes.c_int),
                ("prev_frames", ctypes.c_int), ("mapping", ctypes.c_int), ("width", ctypes.c_int), ("height", ctypes.c_int),
                ("lockstep", ctypes.c_int), ("filter_local_map", ctypes.c_int)]


def load():
    if not os.path.exists(HOST_SO):
        raise RuntimeError("%s not built: run __graft_entry__.build()" % HOST_SO)
    lib = ctypes.CDLL(HOST_SO)
    lib.liodom_host_run_sequence.restype = ctypes.c_int
    return lib


def run_sequence(scans, results_dir="", lockstep=True, width=0, height=0, **kw):
    """-> (poses [n,4,4], nfeats [n], produced). Parameters use the ROS names of the reference."""
    lib = load()
    o = HostOptions(kw.get("min_range", 3.0), kw.get("max_range", 75.0), kw.get("lidar_type", 0), kw.get("scan_lines", 64),
                    kw.get("scan_regions", 8), kw.get("edges_per_region", 10), kw.get("prev_frames", 5), int(kw.get("mapping", 0)),
                    width, height, 1 if lockstep else 0, int(kw.get("filter_local_map", 0)))
    npts = np.array([len(s) for s in scans], np.int32)
    pts = np.ascontiguousarray(np.concatenate(scans)[:, :4], dtype=np.float32)
    poses = np.zeros((len(scans), 16))
    nf = np.zeros(len(scans), np.int32)
    vp = ctypes.c_void_p
    produced = lib.liodom_host_run_sequence(ctypes.byref(o), pts.ctypes.data_as(vp), npts.ctypes.data_as(vp), len(scans),
                                            poses.ctypes.data_as(vp), nf.ctypes.data_as(vp), results_dir.encode())
    return poses.reshape(-1, 4, 4), nf, produced


def last_run_times(n):
    """-> (push_ms [n], pose_ms [n]): wall-clock marks of the last run_sequence call, ms since its start."""
    lib = load()
    lib.liodom_host_last_run_times.restype = ctypes.c_int
    push, pose = np.zeros(n), np.zeros(n)
    m = lib.liodom_host_last_run_times(push.ctypes.data_as(ctypes.c_void_p), pose.ctypes.data_as(ctypes.c_void_p), n)
    return push[:m], pose[:m]


FLOAT32, UINT16, FLOAT64 = 7, 4, 8   # sensor_msgs/PointField datatypes


def run_sequence_msgs(blobs, widths, heights, point_step, fields, row_pad=0, results_dir="", lockstep=True, **kw):
    """Like run_sequence, but every frame is a raw sensor_msgs/PointCloud2 data blob (uint8 array of
    height * (width * point_step + row_pad) bytes) and `fields` = [(name, offset, datatype), ...].
    The façade derives the layout by pcl::fromROSMsg's rules and ships the bytes to the GPU undecoded."""
    lib = load()
    lib.liodom_host_run_sequence_msgs.restype = ctypes.c_int
    o = HostOptions(kw.get("min_range", 3.0), kw.get("max_range", 75.0), kw.get("lidar_type", 0), kw.get("scan_lines", 64),
                    kw.get("scan_regions", 8), kw.get("edges_per_region", 10), kw.get("prev_frames", 5), int(kw.get("mapping", 0)),
                    0, 0, 1 if lockstep else 0, int(kw.get("filter_local_map", 0)))
    data = np.ascontiguousarray(np.concatenate([np.asarray(b, np.uint8).ravel() for b in blobs]))
    w = np.array(widths, np.int32)
    h = np.array(heights, np.int32)
    offs = np.array([f[1] for f in fields], np.int32)
    dts = np.array([f[2] for f in fields], np.int32)
    names = ",".join(f[0] for f in fields).encode()
    n = len(blobs)
    poses = np.zeros((n, 16))
    nf = np.zeros(n, np.int32)
    vp = ctypes.c_void_p
    produced = lib.liodom_host_run_sequence_msgs(ctypes.byref(o), data.ctypes.data_as(vp), w.ctypes.data_as(vp), h.ctypes.data_as(vp), n,
                                                 point_step, row_pad, names, offs.ctypes.data_as(vp), dts.ctypes.data_as(vp), len(fields),
                                                 poses.ctypes.data_as(vp), nf.ctypes.data_as(vp), results_dir.encode())
    return poses.reshape(-1, 4, 4), nf, produced


def make_odometry(pose, prev_odom, laser_to_base, stamp, prev_stamp):
    """LaserOdometer::publishOdom's arithmetic through the façade -> 13 doubles (orientation x,y,z,w,
    position, twist linear, twist angular)."""
    lib = load()
    lib.liodom_host_make_odometry.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    a = [np.ascontiguousarray(m, dtype=np.float64) for m in (pose, prev_odom, laser_to_base)]
    out = np.empty(13)
    lib.liodom_host_make_odometry(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, float(stamp), float(prev_stamp), out.ctypes.data)
    return out
