"""GPU parity: FeatureExtractor path (split, curvature, region selection) through the C ABI
against the oracle — bit-exact (ring ids, ring-major order, smoothness keys, edge lists)."""
import numpy as np
import pytest

import oracle
from liodom_b200 import api, synth
from conftest import get_sequence

pytestmark = pytest.mark.gpu


def _cmp_split(g, o):
    assert g["n_ambiguous"] == 0, "scan has ring-bin decisions within 1e-12 of a boundary"
    assert np.array_equal(g["ring_of_point"], o["ring_of_point"])
    assert np.array_equal(g["offsets"], o["offsets"])
    assert np.array_equal(g["src_index"], o["src_index"])
    assert np.array_equal(g["rings"].view(np.uint32), o["rings"].view(np.uint32))


def _cmp_extract(ctx, op, scan, width=0, height=0):
    sp = oracle.split(op, scan, width, height)
    o0 = oracle.extract(op, sp["rings"], sp["offsets"], sort_mode=0, want_keys=True)
    o1 = oracle.extract(op, sp["rings"], sp["offsets"], sort_mode=1)
    # the reference's std::sort leaves ties unspecified; the data must not depend on it
    assert np.array_equal(o0["idx"], o1["idx"]) and np.array_equal(o0["ring"], o1["ring"])
    g = ctx.extract(scan, width=width, height=height, debug=True)
    assert len(g["edges"]) == len(o0["edges"])
    assert np.array_equal(g["ring"], o0["ring"])
    assert np.array_equal(g["idx"], o0["idx"])
    assert np.array_equal(g["edges"].view(np.uint32), o0["edges"].view(np.uint32))
    nv = sp["offsets"][-1]
    gk = g["keys"][:nv].view(np.uint64)
    ok = o0["keys"].view(np.uint64)
    ev = ~np.isnan(o0["keys"])
    assert np.array_equal(gk[ev], ok[ev]), "smoothness keys differ bitwise"
    return len(o0["edges"])


@pytest.mark.parametrize("sensor,kw", [("hdl64", {}), ("hdl64_firing", {}), ("hdl64_small", {})])
def test_split_velodyne(cuda_lib, sensor, kw):
    scans, _ = get_sequence(sensor, 1000, 2)
    op = oracle.make_params()
    ctx = api.Context(max_points=131072)
    for s in scans:
        _cmp_split(ctx.split(s), oracle.split(op, s))
    ctx.close()


def test_split_pcl_stride_and_invalid_points(cuda_lib):
    """32-byte pcl::PointXYZI records (intensity at +16), NaN/inf and out-of-range points."""
    s = get_sequence("hdl64_small", 1001, 1)[0][0].copy()
    rng = np.random.default_rng(5)
    bad = rng.choice(len(s), 200, replace=False)
    s[bad[:50], 0] = np.nan
    s[bad[50:100], 2] = np.inf
    s[bad[100:150], :3] *= 100.0   # beyond max_range
    s[bad[150:], :3] *= 0.01       # inside min_range
    wide = np.zeros((len(s), 8), np.float32)
    wide[:, :3] = s[:, :3]
    wide[:, 3] = 1.0
    wide[:, 4] = s[:, 3]
    op = oracle.make_params()
    ctx = api.Context(max_points=32768)
    o = oracle.split(op, s)
    g = ctx.split(wide)
    assert np.array_equal(g["ring_of_point"], o["ring_of_point"])
    assert np.array_equal(g["rings"].view(np.uint32), o["rings"].view(np.uint32))
    assert (o["ring_of_point"][bad] == -1).all()
    ctx.close()


@pytest.mark.parametrize("lines", [16, 32])
def test_split_vlp16_hdl32(cuda_lib, lines):
    """Ring formulas for scan_lines 16 / 32 (src/feature_extractor.cc:139-148) on random rays."""
    rng = np.random.default_rng(lines)
    n = 20000
    az = rng.uniform(-np.pi, np.pi, n)
    el = np.deg2rad(rng.uniform(-35, 20, n))
    r = rng.uniform(1, 90, n)
    pts = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el), rng.uniform(0, 1, n)], 1).astype(np.float32)
    op = oracle.make_params(scan_lines=lines)
    ctx = api.Context(scan_lines=lines, max_points=32768)
    g = ctx.split(pts)
    o = oracle.split(op, pts)
    amb = g["n_ambiguous"]
    assert amb == 0
    _cmp_split(g, o)
    ctx.close()


def test_extract_hdl64_c1(cuda_lib):
    scans, _ = get_sequence("hdl64", 1000, 3)
    op = oracle.make_params(prev_frames=15)
    ctx = api.Context(prev_frames=15, max_points=131072)
    for s in scans:
        e = _cmp_extract(ctx, op, s)
        assert 0 < e <= 5632
    ctx.close()


def test_extract_firing_order(cuda_lib):
    scans, _ = get_sequence("hdl64_firing", 1002, 2)
    op = oracle.make_params()
    ctx = api.Context(max_points=131072)
    for s in scans:
        _cmp_extract(ctx, op, s)
    ctx.close()


def test_extract_ouster_c2(cuda_lib):
    scans, _ = get_sequence("os1_128", 1000, 2)
    w, h = synth.sensor_shape("os1_128")
    op = oracle.make_params(lidar_type=1, scan_lines=128)  # launch/liodom_ouster.launch:17-31, scan_lines=128
    ctx = api.Context(lidar_type=1, scan_lines=128, max_points=262144)
    for s in scans:
        assert len(s) == w * h
        _cmp_extract(ctx, op, s, w, h)
    ctx.close()


def test_extract_stress_c3(cuda_lib):
    scans, _ = get_sequence("hdl64", 1003, 2)
    op = oracle.make_params(scan_regions=16, edges_per_region=20, prev_frames=20)
    ctx = api.Context(scan_regions=16, edges_per_region=20, prev_frames=20, max_points=131072)
    for s in scans:
        e = _cmp_extract(ctx, op, s)
        assert e <= 64 * 16 * 21
    ctx.close()


def test_extract_long_rings_1m(cuda_lib):
    """1M-point scan: rings (15,625 pts) exceed the shared-memory ring capacity."""
    s = get_sequence("hdl64_1m", 1000, 1)[0][0]
    op = oracle.make_params()
    ctx = api.Context(max_points=1 << 20)
    _cmp_extract(ctx, op, s)
    op2 = oracle.make_params(scan_regions=64)
    ctx2 = api.Context(scan_regions=64, max_points=1 << 20)
    _cmp_extract(ctx2, op2, s)
    ctx.close()
    ctx2.close()


def test_extract_edge_cases(cuda_lib):
    op = oracle.make_params()
    ctx = api.Context(max_points=32768)
    # empty scan
    g = ctx.extract(np.zeros((0, 4), np.float32))
    assert len(g) == 0
    # a scan whose rings are all shorter than min_points_per_scan_ (=90)
    s = get_sequence("hdl64_small", 1000, 1)[0][0]
    sp = oracle.split(op, s)
    keep = np.concatenate([sp["src_index"][sp["offsets"][r]:sp["offsets"][r + 1]][:89] for r in range(64)])
    short = s[np.sort(keep)]
    assert len(ctx.extract(short)) == 0
    # rings exactly at the threshold (90 points) and ragged lengths
    keep = np.concatenate([sp["src_index"][sp["offsets"][r]:sp["offsets"][r + 1]][:90 + 3 * r] for r in range(64)])
    ragged = s[np.sort(keep)]
    _cmp_extract(ctx, op, ragged)
    ctx.close()


def test_extract_capacity_and_errors(cuda_lib):
    ctx = api.Context(max_points=2048)
    with pytest.raises(api.LiodomError):
        ctx.extract(np.ones((5000, 4), np.float32))
    ctx.close()
    with pytest.raises(api.LiodomError):
        api.Context(scan_lines=48)   # "Invalid scan lines" (src/feature_extractor.cc:150)
    with pytest.raises(api.LiodomError):
        api.Context(lidar_type=2)    # "Incorrect Lidar type" (:177)


@pytest.mark.parametrize("lines", [64, 32, 16])
def test_split_points_near_bin_and_range_boundaries(cuda_lib, lines):
    """The ring split screens every decision in FP32 and falls back to the reference's double arithmetic
    when a threshold is within reach of the float error.  Points placed at 1e-7 .. 1e-2 degrees from every
    bin edge and 1e-6 .. 1e-2 m from the range gates must get the oracle's ring ids."""
    rng = np.random.default_rng(100 + lines)
    if lines == 64:
        edges = [2.0 - (k - 0.5) / 3.0 for k in range(0, 34)] + [-8.83 - (k - 0.5) / 2.0 for k in range(0, 33)] + [2.0, -8.83, -24.33]
    elif lines == 32:
        edges = [k * 4.0 / 3.0 - 92.0 / 3.0 for k in range(-1, 34)]
    else:
        edges = [2.0 * (k - 0.5) - 15.0 for k in range(-1, 18)]
    pts = []
    for e in edges:
        for d in (1e-2, 1e-3, 5e-4, 2e-4, 1e-5, 1e-7):
            for sgn in (-1.0, 1.0):
                ang = np.deg2rad(e + sgn * d)
                az = rng.uniform(0, 2 * np.pi)
                r = rng.uniform(5.0, 60.0)
                pts.append([r * np.cos(az), r * np.sin(az), r * np.tan(ang), 0.5])
    for gate in (3.0, 75.0):
        for d in (1e-2, 1e-3, 1e-4, 1e-5, 1e-6):
            for sgn in (-1.0, 1.0):
                az = rng.uniform(0, 2 * np.pi)
                r = gate + sgn * d
                pts.append([r * np.cos(az), r * np.sin(az), r * np.tan(np.deg2rad(-3.1)), 0.25])
    s = np.array(pts, np.float32)
    op = oracle.make_params(scan_lines=lines)
    ctx = api.Context(scan_lines=lines, max_points=4096)
    g = ctx.split(s)
    o = oracle.split(op, s)
    # device atan (<= 2 ulp) and glibc atan (<= 1 ulp) may differ within 1e-12 of an edge: those are counted, not compared
    assert g["n_ambiguous"] <= 4
    diff = np.nonzero(g["ring_of_point"] != o["ring_of_point"])[0]
    assert len(diff) <= g["n_ambiguous"], (len(diff), g["n_ambiguous"])
    assert (g["ring_of_point"] >= 0).sum() > len(s) // 2
    ctx.close()


def test_large_region_configs_and_limits(cuda_lib):
    """k_compact / k_extract raise their dynamic shared-memory limit per device at context creation: a 64 x 256
    region grid (66 KB of prefix-sum scratch, above the 48 KB default) must run; a configuration whose
    selection scratch cannot fit the 227 KB of an SM is refused at creation, not at launch."""
    s = get_sequence("hdl64", 1000, 1)[0][0]
    kw = dict(scan_regions=200, edges_per_region=1)
    op = oracle.make_params(**kw)
    ctx = api.Context(max_points=131072, **kw)
    _cmp_extract(ctx, op, s)
    ctx.close()
    with pytest.raises(api.LiodomError):
        api.Context(max_points=131072, scan_regions=256, edges_per_region=255)
    with pytest.raises(api.LiodomError):      # window + received map beyond the voxel hash's 2^20-point limit
        api.Context(max_points=131072, mapping=1, max_received_map=1 << 20)


def test_two_contexts_on_two_devices(cuda_lib):
    """One process, one context per GPU: each device needs its own shared-memory opt-in."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = get_sequence("hdl64", 1000, 1)[0][0]
    op = oracle.make_params()
    sp = oracle.split(op, s)
    oe = oracle.extract(op, sp["rings"], sp["offsets"])["edges"]
    for dev in (1, 0):
        ctx = api.Context(max_points=131072, device=dev)
        g = ctx.extract(s)
        assert np.array_equal(g.view(np.uint32), oe.view(np.uint32)), dev
        ctx.close()
