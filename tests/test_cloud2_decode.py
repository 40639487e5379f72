"""sensor_msgs/PointCloud2 decode (pcl::fromROSMsg, src/liodom_node.cc:43-44): the oracle's restatement
against NumPy structured arrays (CPU), and the device path that reads the fields in place from the
raw message bytes against the oracle (GPU): edges bit-exact for every wire layout."""
import numpy as np
import pytest

import oracle
from conftest import get_sequence


def _blob(scan, dtype, width=None, height=1, row_pad=0, rng=None):
    """Pack an [n,4] scan into PointCloud2 bytes with structured dtype `dtype` (+ junk in the other
    fields and in the row padding)."""
    n = len(scan)
    rng = rng or np.random.default_rng(0)
    rec = np.zeros(n, dtype)
    raw = rec.view(np.uint8).reshape(n, dtype.itemsize)
    raw[:] = rng.integers(0, 256, raw.shape, dtype=np.uint8)   # junk everywhere first
    rec = raw.view(dtype).reshape(n)
    rec["x"], rec["y"], rec["z"] = scan[:, 0], scan[:, 1], scan[:, 2]
    if "intensity" in dtype.names:
        rec["intensity"] = scan[:, 3]
    raw = rec.view(np.uint8).reshape(n, dtype.itemsize)
    if row_pad == 0:
        return raw.reshape(-1).copy()
    width = width or n
    rows = raw.reshape(height, width * dtype.itemsize)
    pad = rng.integers(0, 256, (height, row_pad), dtype=np.uint8)
    return np.concatenate([rows, pad], 1).reshape(-1).copy()


LAYOUTS = {
    # velodyne_pointcloud PointXYZIRT: 22-byte points, not a multiple of 4
    "velodyne_xyzirt22": np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                                   "offsets": [0, 4, 8, 12, 16, 18], "itemsize": 22}),
    # pcl::PointXYZI as published by PCL nodes: intensity at +16, 32-byte points
    "pcl_xyzi32": np.dtype({"names": ["x", "y", "z", "intensity"], "formats": ["<f4"] * 4, "offsets": [0, 4, 8, 16], "itemsize": 32}),
    # ouster_ros Point: x,y,z at 0..8, intensity at +16, 48-byte points
    "ouster48": np.dtype({"names": ["x", "y", "z", "intensity", "t", "reflectivity", "ring", "ambient", "range"],
                          "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2", "<u1", "<u2", "<u4"],
                          "offsets": [0, 4, 8, 16, 20, 24, 26, 28, 32], "itemsize": 48}),
    # fields in another order, odd offsets, no intensity
    "odd_no_intensity": np.dtype({"names": ["tag", "z", "x", "y"], "formats": ["<u1", "<f4", "<f4", "<f4"], "offsets": [0, 1, 7, 13], "itemsize": 19}),
}


def _offsets(dt):
    f = dt.fields
    return f["x"][1], f["y"][1], f["z"][1], (f["intensity"][1] if "intensity" in f else -1)


@pytest.mark.parametrize("name", sorted(LAYOUTS))
def test_oracle_decode_matches_numpy(name):
    dt = LAYOUTS[name]
    rng = np.random.default_rng(3)
    scan = rng.normal(size=(12 * 7, 4)).astype(np.float32)
    for row_pad in (0, 5):
        blob = _blob(scan, dt, width=12, height=7, row_pad=row_pad, rng=rng)
        ox, oy, oz, oi = _offsets(dt)
        out = oracle.decode_cloud2(blob, 12, 7, dt.itemsize, 12 * dt.itemsize + row_pad, ox, oy, oz, oi)
        want = scan.copy()
        if oi < 0:
            want[:, 3] = 0
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(LAYOUTS))
def test_gpu_extract_from_raw_message(cuda_lib, name):
    from liodom_b200 import api
    dt = LAYOUTS[name]
    scan = get_sequence("hdl64", 1000, 1)[0][0]
    ox, oy, oz, oi = _offsets(dt)
    blob = _blob(scan, dt)
    op = oracle.make_params()
    dec = oracle.decode_cloud2(blob, len(scan), 1, dt.itemsize, len(scan) * dt.itemsize, ox, oy, oz, oi)
    sp = oracle.split(op, dec)
    o = oracle.extract(op, sp["rings"], sp["offsets"])
    ctx = api.Context(max_points=131072)
    lay = api.CloudLayout(dt.itemsize, 0, ox, oy, oz, oi, 0)
    g = ctx.extract_layout(blob, len(scan), lay)
    assert len(o["edges"]) > 1000
    assert np.array_equal(g["ring"], o["ring"]) and np.array_equal(g["idx"], o["idx"])
    assert np.array_equal(g["edges"].view(np.uint32), o["edges"].view(np.uint32))
    ctx.close()


@pytest.mark.gpu
def test_gpu_organised_message_with_row_padding(cuda_lib):
    """OS1-128-shaped organised cloud, 48-byte ouster points, rows padded by 16 bytes."""
    from liodom_b200 import api, synth
    w, h = synth.sensor_shape("os1_128")
    scan = get_sequence("os1_128", 1000, 1)[0][0]
    dt = LAYOUTS["ouster48"]
    ox, oy, oz, oi = _offsets(dt)
    blob = _blob(scan, dt, width=w, height=h, row_pad=16)
    op = oracle.make_params(lidar_type=1, scan_lines=128)
    dec = oracle.decode_cloud2(blob, w, h, dt.itemsize, w * dt.itemsize + 16, ox, oy, oz, oi)
    assert np.array_equal(dec.view(np.uint32), scan.view(np.uint32))
    sp = oracle.split(op, dec, w, h)
    o = oracle.extract(op, sp["rings"], sp["offsets"])
    ctx = api.Context(lidar_type=1, scan_lines=128, max_points=262144)
    lay = api.CloudLayout(dt.itemsize, w * dt.itemsize + 16, ox, oy, oz, oi, 0)
    g = ctx.extract_layout(blob, w * h, lay, width=w, height=h)
    assert len(o["edges"]) > 1000
    assert np.array_equal(g["edges"].view(np.uint32), o["edges"].view(np.uint32))
    ctx.close()


@pytest.mark.gpu
def test_gpu_layout_errors(cuda_lib):
    from liodom_b200 import api
    ctx = api.Context(max_points=4096)
    blob = np.zeros(22 * 10, np.uint8)
    with pytest.raises(api.LiodomError):   # big-endian: pcl::fromROSMsg would misread it
        ctx.extract_layout(blob, 10, api.CloudLayout(22, 0, 0, 4, 8, 12, 1))
    with pytest.raises(api.LiodomError):   # field outside the point
        ctx.extract_layout(blob, 10, api.CloudLayout(22, 0, 0, 4, 20, 12, 0))
    ctx.close()


@pytest.mark.gpu
def test_gpu_batched_messages_and_facade(cuda_lib):
    """Whole path fed with raw 22-byte velodyne messages: the batched C-ABI call and the façade's
    lidarClb route (fromROSMsgDeferred -> SharedData -> FeatureExtractor) give the poses of the
    decoded-array route."""
    from liodom_b200 import api, host_api
    scans, _ = get_sequence("hdl64_small", 1000, 6)
    dt = LAYOUTS["velodyne_xyzirt22"]
    ox, oy, oz, oi = _offsets(dt)
    blobs = [_blob(s, dt) for s in scans]
    ref = api.Context(prev_frames=5, max_points=32768)
    ctx = api.Context(prev_frames=5, max_points=32768)
    lay = api.CloudLayout(dt.itemsize, 0, ox, oy, oz, oi, 0)
    want = []
    for s, b in zip(scans, blobs):
        ref.scan_batch([s])
        pr, ner = ref.results()
        ctx.scan_batch_layout([b], [len(s)], lay)
        pg, neg = ctx.results()
        assert neg[0] == ner[0] and np.array_equal(pg, pr)
        want.append(pr[0].copy())
    ref.close()
    ctx.close()
    fields = [("x", 0, host_api.FLOAT32), ("y", 4, host_api.FLOAT32), ("z", 8, host_api.FLOAT32), ("intensity", 12, host_api.FLOAT32),
              ("ring", 16, host_api.UINT16), ("time", 18, host_api.FLOAT32)]
    poses, nf, produced = host_api.run_sequence_msgs(blobs, [len(s) for s in scans], [1] * len(scans), 22, fields, prev_frames=5)
    assert produced == len(scans)
    assert np.array_equal(poses, np.stack(want))
