"""CPU tests: the C++ oracle of the FeatureExtractor path against an independent NumPy
restatement and the known-answer invariants read off the reference."""
import numpy as np
import pytest

import oracle
import np_ref
from conftest import get_sequence


def _ring_line(n, step=0.05, y=10.0, z=-1.0):
    """A straight, evenly spaced ring segment: zero curvature everywhere."""
    p = np.zeros((n, 4), np.float32)
    p[:, 0] = np.arange(n) * step - n * step / 2
    p[:, 1] = y
    p[:, 2] = z
    return p


def _extract_single_ring(ring, **kw):
    p = oracle.make_params(**kw)
    off = np.zeros(p.scan_lines + 1, np.int32)
    off[1:] = len(ring)
    return oracle.extract(p, ring, off, sort_mode=0, want_keys=True)


def test_hdl64_beam_k_maps_to_ring_k():
    """Nominal HDL-64 beam k (2 - k/3 deg, k<32; -8.83 - (k-32)/2 deg otherwise) -> ring k
    (src/feature_extractor.cc:128-138). Beams 0 and 63 sit ON the reject limits (2 / -24.33 deg),
    so they are nudged inwards by 1e-3 deg."""
    p = oracle.make_params()
    pts = []
    for k in range(64):
        el = (2.0 - k / 3.0) if k < 32 else (-8.83 - (k - 32) / 2.0)
        el = el - 1e-3 if k == 0 else (el + 1e-3 if k == 63 else el)
        for az in (0.1, 2.0, 4.5):
            r = 20.0
            pts.append([r * np.cos(np.deg2rad(el)) * np.cos(az), r * np.cos(np.deg2rad(el)) * np.sin(az), r * np.sin(np.deg2rad(el)), 0])
    pts = np.array(pts, np.float32)
    o = oracle.split(p, pts)
    assert np.array_equal(o["ring_of_point"], np.repeat(np.arange(64), 3))
    assert [np_ref.ring_hdl64(*q[:3]) for q in pts] == list(o["ring_of_point"])


def test_split_matches_numpy_and_is_stable():
    s = get_sequence("hdl64_small", 1000, 1)[0][0]
    p = oracle.make_params()
    o = oracle.split(p, s)
    ref = np.array([np_ref.ring_hdl64(*q[:3]) for q in s[:4000]])
    assert np.array_equal(o["ring_of_point"][:4000], ref)
    # ring-major, order within a ring = input order
    for r in range(64):
        src = o["src_index"][o["offsets"][r]:o["offsets"][r + 1]]
        assert (np.diff(src) > 0).all()
        assert (o["ring_of_point"][src] == r).all()
    assert np.array_equal(o["rings"], s[o["src_index"]])


def test_range_filter_is_xy_only():
    """isValidPoint uses sqrt(x^2+y^2), not the 3-D range (src/feature_extractor.cc:96-97)."""
    p = oracle.make_params()
    pts = np.array([[2.9, 0, -0.5, 0],      # xy 2.9 < 3 -> rejected even though 3-D range is 2.94
                    [3.0, 0, -0.5, 0],      # on the limit: kept (rejects only < min)
                    [0, 75.0, -3.0, 0],     # on the limit: kept
                    [0, 75.01, -3.0, 0],    # rejected
                    [1.0, 1.0, -60.0, 0],   # 3-D range 60 m but xy 1.41 -> rejected
                    [np.nan, 5, 0, 0], [5, np.inf, 0, 0], [5, 5, -np.inf, 0]], np.float32)
    o = oracle.split(p, pts)
    assert list(o["ring_of_point"] >= 0) == [False, True, True, False, False, False, False, False]


def test_vlp16_hdl32_and_bad_configs():
    rng = np.random.default_rng(0)
    el = np.deg2rad(rng.uniform(-30, 15, 2000))
    az = rng.uniform(-3, 3, 2000)
    pts = np.stack([20 * np.cos(el) * np.cos(az), 20 * np.cos(el) * np.sin(az), 20 * np.sin(el), 0 * el], 1).astype(np.float32)
    ang = np.arctan(pts[:, 2].astype(np.float64) / np.sqrt(pts[:, 0].astype(np.float64) ** 2 + pts[:, 1].astype(np.float64) ** 2)) * 180 / np.pi
    o16 = oracle.split(oracle.make_params(scan_lines=16), pts)["ring_of_point"]
    e16 = np.array([int((a + 15) / 2 + 0.5) for a in ang])
    e16[(e16 > 15) | (e16 < 0)] = -1
    assert np.array_equal(o16, e16)
    o32 = oracle.split(oracle.make_params(scan_lines=32), pts)["ring_of_point"]
    e32 = np.array([int((a + 92.0 / 3.0) * 3.0 / 4.0) for a in ang])
    e32[(e32 > 31) | (e32 < 0)] = -1
    assert np.array_equal(o32, e32)
    # invalid scan_lines / lidar_type: logged by the reference, nothing emitted
    assert oracle.split(oracle.make_params(scan_lines=48), pts)["status"] < 0
    assert oracle.split(oracle.make_params(lidar_type=2), pts)["status"] < 0


def test_ouster_rows_are_rings():
    """lidar_type 1: ring = row of the organised cloud, invalid points dropped (:160-175)."""
    h, w = 8, 64
    rng = np.random.default_rng(1)
    pts = rng.uniform(-20, 20, (h * w, 4)).astype(np.float32)
    pts[::5] = 0.0   # no-return slots
    p = oracle.make_params(lidar_type=1, scan_lines=8)
    o = oracle.split(p, pts, w, h)
    d = np.sqrt(pts[:, 0].astype(np.float64) ** 2 + pts[:, 1].astype(np.float64) ** 2)
    exp = np.where((d >= 3.0) & (d <= 75.0), np.arange(h * w) // w, -1)
    assert np.array_equal(o["ring_of_point"], exp)
    # a cloud taller than scan_lines overruns `scans` in the reference: reported as an error
    assert oracle.split(oracle.make_params(lidar_type=1, scan_lines=4), pts, w, h)["status"] < 0


def test_curvature_and_selection_match_numpy():
    s = get_sequence("hdl64_small", 1000, 1)[0][0]
    p = oracle.make_params()
    sp = oracle.split(p, s)
    o = oracle.extract(p, sp["rings"], sp["offsets"], sort_mode=0, want_keys=True)
    o1 = oracle.extract(p, sp["rings"], sp["offsets"], sort_mode=1)
    assert np.array_equal(o["idx"], o1["idx"]) and np.array_equal(o["ring"], o1["ring"])
    for r in (0, 7, 31, 32, 50, 63):
        ring = sp["rings"][sp["offsets"][r]:sp["offsets"][r + 1]]
        k = np_ref.curvature(ring)
        ok = o["keys"][sp["offsets"][r]:sp["offsets"][r + 1]]
        m = ~np.isnan(k)
        assert np.array_equal(k[m].view(np.uint64), ok[m].view(np.uint64))
        assert np.isnan(ok[~m]).all()
        assert np_ref.select_ring(ring) == list(o["idx"][o["ring"] == r])
    # edges are copies of the ring points
    for e, r, i in zip(o["edges"][:200], o["ring"][:200], o["idx"][:200]):
        assert np.array_equal(e, sp["rings"][sp["offsets"][r] + i])


def test_region_yields_at_most_epr_plus_one():
    """`picked_edges > edges_per_region` (src/feature_extractor.cc:270) lets epr+1 edges through."""
    n = 8 * 10 + 10 + 2000
    ring = _ring_line(n, step=0.3)           # gaps^2 = 0.09 > 0.05: no neighbour suppression
    rng = np.random.default_rng(2)
    ring[:, 2] += rng.normal(0, 0.2, n).astype(np.float32)   # every point is "sharp"
    o = _extract_single_ring(ring)
    counts = np.bincount((o["idx"][o["ring"] == 0] - 5) // ((n - 10) // 8), minlength=8)[:8]
    assert (counts == 11).all()


def test_threshold_and_break_on_first_smooth_item():
    """No edge below 0.1; the walk stops at the first un-picked item below 0.1 (:270)."""
    n = 500
    ring = _ring_line(n, step=0.3)
    ring[100, 2] += 1.0     # one spike -> keys at 95..105 large
    o = _extract_single_ring(ring)
    keys = o["keys"][:n]
    sel = o["idx"][o["ring"] == 0]
    assert len(sel) > 0 and (keys[sel] >= 0.1).all()
    assert set(sel) <= set(range(95, 106))
    # flat ring: nothing at all
    assert len(_extract_single_ring(_ring_line(n))["edges"]) == 0


def test_suppression_window_and_gap_break():
    """After picking p, p+-1..5 are suppressed until a consecutive gap^2 > 0.05 (:280-310)."""
    n = 400
    ring = _ring_line(n, step=0.05)
    ring[:, 2] += np.abs(np.arange(n) - 200).astype(np.float32) * 0.05   # a "V": corner at 200, gaps^2 = 0.005
    o = _extract_single_ring(ring)
    sel = sorted(o["idx"][o["ring"] == 0])
    assert 200 in sel
    assert not any(195 <= i <= 205 and i != 200 for i in sel)            # the full +-5 window is suppressed
    # with a wide gap right after the corner the forward suppression stops immediately
    ring2 = ring.copy()
    ring2[201:, 0] += 0.5                     # gap^2 between 200 and 201 > 0.05
    o2 = _extract_single_ring(ring2)
    sel2 = sorted(o2["idx"][o2["ring"] == 0])
    near = [i for i in sel2 if 190 <= i <= 210]
    assert len(near) >= 2 and np.diff(near).min() <= 5     # two picks closer than the window: only the gap allows it
    assert sel2 == sorted(np_ref.select_ring(ring2))


def test_suppression_persists_across_regions():
    """picked_ is shared by the regions of a ring (:230, :268, :293): a pick at the end of region r
    suppresses the first points of region r+1."""
    n = 8 * 100 + 10
    ring = _ring_line(n, step=0.05)
    sector = (n - 10) // 8
    last = sector - 1 + 5                     # last ring index of region 0
    ring[:, 2] += np.abs(np.arange(n) - last).astype(np.float32) * 0.05   # "V" corner at the region boundary
    o = _extract_single_ring(ring)
    keys = o["keys"][:n]
    assert keys[last + 1] >= 0.1 and keys[last + 2] >= 0.1    # region 1 would pick these on its own ...
    sel = list(o["idx"][o["ring"] == 0])
    assert sel == [last]                                       # ... but region 0's pick suppressed them
    assert sel == np_ref.select_ring(ring)


def test_short_rings_contribute_nothing_and_last_region_absorbs_remainder():
    """min_points_per_scan_ = regions*epr + 10 (src/params.cc:63, feature_extractor.cc:188);
    last region ends at total_points (:244-247)."""
    rng = np.random.default_rng(3)
    short = _ring_line(89, step=0.3)
    short[:, 2] += rng.normal(0, 0.3, 89).astype(np.float32)
    assert len(_extract_single_ring(short)["edges"]) == 0
    ok = _ring_line(90, step=0.3)
    ok[:, 2] += rng.normal(0, 0.3, 90).astype(np.float32)
    assert len(_extract_single_ring(ok)["edges"]) > 0
    n = 10 + 8 * 40 + 7                        # remainder 7 goes to region 7
    ring = _ring_line(n, step=0.3)
    ring[n - 6, 2] += 1.0                      # last evaluated index n-6 lies in the remainder
    sel = _extract_single_ring(ring)["idx"]
    assert len(sel) > 0 and sel.max() >= 5 + 8 * 40


def test_golden_extract_fixture():
    """Pins the oracle against the committed fixture (tests/golden/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "extract_hdl64_small.npz"))
    s = get_sequence("hdl64_small", int(g["seed"]), 1)[0][0]
    assert np.array_equal(s, g["scan"])
    p = oracle.make_params()
    sp = oracle.split(p, s)
    o = oracle.extract(p, sp["rings"], sp["offsets"], want_keys=True)
    assert np.array_equal(sp["offsets"], g["offsets"])
    assert np.array_equal(o["ring"], g["edge_ring"]) and np.array_equal(o["idx"], g["edge_idx"])
    assert np.array_equal(o["edges"], g["edges"])
    assert np.nansum(o["keys"]).view(np.uint64) == g["keys_nansum"].view(np.uint64)
