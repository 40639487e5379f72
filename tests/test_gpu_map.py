"""GPU parity: liodom::Map (updateMap / getMap / getLocalMap) through the C ABI against the
oracle: exact cell keys in creation order, exact per-cell counts, and — because both sides
accumulate each voxel as "old centroid, then new points in arrival order" — bit-exact centroids."""
import numpy as np
import pytest

import oracle
from liodom_b200 import api
from conftest import get_sequence

pytestmark = pytest.mark.gpu


def _same(gm, om, T=None, cells=(2, 1)):
    gk, gc = gm.cells()
    ok, oc = om.cells()
    assert np.array_equal(gk, ok), "cell keys / creation order differ"
    assert np.array_equal(gc, oc), "per-cell counts differ"
    g, o = gm.get_map(), om.get_map()
    assert g.shape == o.shape
    assert np.array_equal(g.view(np.uint32), o.view(np.uint32)), "centroids differ bitwise"
    if T is not None:
        gl, ol = gm.get_local_map(T, *cells), om.get_local_map(T, *cells)
        assert np.array_equal(gl.view(np.uint32), ol.view(np.uint32))


@pytest.mark.parametrize("xy,z,cxy,cz", [(40.0, 50.0, 2, 1), (30.0, 35.0, 3, 2), (20.0, 25.0, 2, 1)])
def test_map_random_clouds(cuda_lib, xy, z, cxy, cz):
    rng = np.random.default_rng(int(xy))
    gm, om = api.Map(xy, z, 0.4, max_points=1 << 18), oracle.Map(xy, z, 0.4)
    T = np.eye(4)
    from scipy.spatial.transform import Rotation
    for f in range(12):
        n = int(rng.integers(1, 3000))
        pts = (rng.normal(size=(n, 4)) * [30, 30, 2, 1]).astype(np.float32)
        T[:3, :3] = Rotation.from_rotvec([0.01 * f, -0.02, 0.1 * f]).as_matrix()
        T[:3, 3] = [4.0 * f, -1.5 * f, 0.05 * f]
        gm.update(pts, T)
        om.update(pts, T)
        _same(gm, om, T, (cxy, cz))
    assert gm.size()[0] == len(om.get_map())
    gm.close()


def test_map_reaveraging_and_duplicates(cuda_lib):
    """Old centroids count as single points; identical points and repeated clouds."""
    gm, om = api.Map(40.0, 50.0, 0.4, max_points=1 << 16), oracle.Map(40.0, 50.0, 0.4)
    a = np.array([[1.00, 1.0, 1.0, 0.2], [1.10, 1.0, 1.0, 0.4], [1.10, 1.0, 1.0, 0.4], [-0.01, 0.0, 0.0, 1.0]], np.float32)
    for _ in range(4):
        gm.update(a, np.eye(4))
        om.update(a, np.eye(4))
        _same(gm, om, np.eye(4))
    gm.update(np.zeros((0, 4), np.float32), np.eye(4))     # empty cloud: no-op
    _same(gm, om)
    gm.close()


def test_map_replay_edge_clouds_c4(cuda_lib):
    """C4-style replay: edge clouds of a synthetic sequence + ground-truth poses through
    updateMap -> getLocalMap every frame (src/liodom_mapping_node.cc:45-90)."""
    scans, gt = get_sequence("hdl64_small", 1000, 25)
    p = oracle.make_params()
    gm, om = api.Map(20.0, 25.0, 0.4, max_points=1 << 19), oracle.Map(20.0, 25.0, 0.4)   # launch/liodom_mapping.launch:15-19
    for f, s in enumerate(scans):
        sp = oracle.split(p, s)
        edges = oracle.extract(p, sp["rings"], sp["offsets"])["edges"]
        T = np.linalg.inv(gt[0]) @ gt[f]
        gm.update(edges, T)
        om.update(edges, T)
        gl, ol = gm.get_local_map(T, 2, 1), om.get_local_map(T, 2, 1)
        assert np.array_equal(gl.view(np.uint32), ol.view(np.uint32)), f
    _same(gm, om, T)
    n, c = gm.size()
    assert n > 10000 and c >= 9
    gm.close()


def test_map_errors(cuda_lib):
    with pytest.raises(api.LiodomError):
        api.Map(0.5, 50.0, 0.4)                      # cell sizes below 1 m
    with pytest.raises(api.LiodomError):
        api.Map(40.0, 50.0, 0.01)                    # more than ~1000 voxels per cell axis
    gm = api.Map(40.0, 50.0, 0.4, max_points=1024)
    with pytest.raises(api.LiodomError):
        gm.update(np.random.default_rng(0).normal(size=(5000, 4)).astype(np.float32) * 30, np.eye(4))
    gm.close()


def test_mapping_feedback_loop(cuda_lib):
    """mapping:=true (launch/liodom.launch:39-57): per frame the odometry pose feeds Map::updateMap,
    Map::getLocalMap feeds the next frame's kNN target (src/laser_odometry.cc:276-278,312-314).  GPU:
    the local map goes map -> odometry on the device; oracle: the same loop on the CPU.  Teacher-forced
    per frame (the oracle's state is loaded into the GPU lane), so poses must agree to 1e-4 m / 1e-5 rad
    and both the window + received-map sizes and the maps must match."""
    from conftest import pose_err
    scans, _ = get_sequence("hdl64_small", 1000, 8)
    op = oracle.make_params(prev_frames=5, mapping=1)
    ctx = api.Context(prev_frames=5, mapping=1, max_points=32768, max_received_map=1 << 18)
    odo = oracle.Odometer(op)
    gm, om = api.Map(30.0, 35.0, 0.4, max_points=1 << 19), oracle.Map(30.0, 35.0, 0.4)    # launch/liodom.launch:46-50
    sizes = []
    for f, s in enumerate(scans):
        sp = oracle.split(op, s)
        edges = oracle.extract(op, sp["rings"], sp["offsets"])["edges"]
        if f > 0:
            w, nf = odo.window()
            ctx.lmap_clear()
            pos = 0
            for n in sizes:
                ctx.lmap_add(w[pos:pos + n])
                pos += n
            o_odom, o_prev = odo.get_pose()
            ctx.set_pose(o_odom, o_prev)
        opose, od = odo.process(edges)
        gpose, gd = ctx.register(edges)
        if f > 0:
            assert gd.n_map[0] == od.n_map[0] and gd.n_map[0] > sum(sizes)      # window + received map
        dt, dr = pose_err(gpose, opose)
        assert dt < 1e-4 and dr < 1e-5, (f, dt, dr)
        sizes.append(len(edges))
        if len(sizes) > 5:
            sizes.pop(0)
        # mapping process: both sides integrate the ORACLE pose so that the maps stay comparable
        gm.update(edges, opose)
        om.update(edges, opose)
        loc = om.get_local_map(opose, 3, 2)
        odo.set_received_map(loc)
        n = ctx.set_received_map_from(gm, opose, 3, 2)
        assert n == len(loc)
    _same(gm, om, opose, (3, 2))
    ctx.close()
    gm.close()
