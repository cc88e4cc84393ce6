"""GPU parity: exact kNN (grid and exhaustive kernels) vs the CPU oracle, through the C ABI.

Bar: neighbour indices AND squared-distance bit patterns identical (integer/bit-exact).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.int32)


def _check(got, want):
    gi, gd, gf = got
    wi, wd, wf = want
    assert (gf == wf).all(), f"found mismatch at {np.nonzero(gf != wf)[0][:5]}"
    bad = np.nonzero((gi != wi).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} queries differ, first {bad[:5]}: got {gi[bad[:2]]} want {wi[bad[:2]]}"
    assert (_bits(gd) == _bits(wd)).all()


@pytest.mark.parametrize("k", [1, 2, 5, 8, 16])
@pytest.mark.parametrize("max_dist", [0.0, 2.0])
def test_grid_knn_matches_oracle(pkg, O, handle, k, max_dist):
    src, tgt, _ = pkg.synth.frame_pair(1001, 1024)
    handle.map_build(tgt)
    _check(handle.map_knn(src, k, max_dist), O.knn(tgt, src, k, max_dist))


@pytest.mark.parametrize("k", [1, 5])
def test_brute_knn_matches_oracle(pkg, O, handle, k):
    src, tgt, _ = pkg.synth.frame_pair(11, 700, 1900)
    handle.map_build(tgt)
    _check(handle.map_knn_brute(src, k, 0.0), O.knn(tgt, src, k, 0.0))
    _check(handle.map_knn_brute(src, k, 1.5), O.knn(tgt, src, k, 1.5))


def test_scan_to_map_shape(pkg, O, handle):
    scan, mp, _ = pkg.synth.scan_to_map(1002, 4096, 200000)
    handle.map_build(mp)
    got = handle.map_knn(scan[:1024], 5, 2.0)
    _check(got, O.knn(mp, scan[:1024], 5, 2.0))
    # full query set: grid kernel vs exhaustive kernel on the device (size-independent self-consistency)
    _check(handle.map_knn(scan, 5, 2.0), handle.map_knn_brute(scan, 5, 2.0))
    _check(handle.map_knn(scan, 5, 0.0), handle.map_knn_brute(scan, 5, 0.0))


def test_far_queries_and_cell_sizes(pkg, O, handle):
    """queries far outside the map (ungated PCL default) and awkward cell sizes stay exact"""
    rng = np.random.default_rng(5)
    _, tgt, _ = pkg.synth.frame_pair(21, 16, 3000)
    q = np.zeros((256, 4), np.float32)
    q[:, :3] = rng.uniform(-500, 500, (256, 3))
    q[:64, :3] = rng.uniform(-90, 90, (64, 3))
    want1, want5 = O.knn(tgt, q, 1, 0.0), O.knn(tgt, q, 5, 0.0)
    want5g = O.knn(tgt, q, 5, 7.5)
    for cell in (0.0, 0.25, 3.0, 50.0, 1000.0):
        handle.map_build(tgt, cell)
        _check(handle.map_knn(q, 1, 0.0), want1)
        _check(handle.map_knn(q, 5, 0.0), want5)
        _check(handle.map_knn(q, 5, 7.5), want5g)


def test_edge_cases(pkg, O, handle):
    rng = np.random.default_rng(3)
    # fewer points than k, a single point, duplicates (ties -> lowest index)
    tgt = np.zeros((3, 4), np.float32)
    tgt[:, :3] = rng.uniform(-1, 1, (3, 3))
    q = np.zeros((10, 4), np.float32)
    q[:, :3] = rng.uniform(-1, 1, (10, 3))
    handle.map_build(tgt)
    _check(handle.map_knn(q, 5, 0.0), O.knn(tgt, q, 5, 0.0))
    handle.map_build(tgt[:1])
    _check(handle.map_knn(q, 2, 0.0), O.knn(tgt[:1], q, 2, 0.0))
    dup = np.repeat(tgt, 4, axis=0)
    handle.map_build(dup)
    _check(handle.map_knn(q, 5, 0.0), O.knn(dup, q, 5, 0.0))
    _check(handle.map_knn_brute(q, 5, 0.0), O.knn(dup, q, 5, 0.0))
    # planar and collinear clouds (degenerate bounding boxes)
    flat = np.zeros((500, 4), np.float32)
    flat[:, :2] = rng.uniform(-10, 10, (500, 2))
    handle.map_build(flat)
    q2 = np.zeros((64, 4), np.float32)
    q2[:, :3] = rng.uniform(-12, 12, (64, 3))
    _check(handle.map_knn(q2, 5, 0.0), O.knn(flat, q2, 5, 0.0))
    line = np.zeros((300, 4), np.float32)
    line[:, 0] = rng.uniform(-10, 10, 300)
    handle.map_build(line)
    _check(handle.map_knn(q2, 3, 0.0), O.knn(line, q2, 3, 0.0))
    # empty query set
    idx, d2, found = handle.map_knn(np.zeros((0, 4), np.float32), 5, 0.0)
    assert idx.shape == (0, 5)


def test_add_points_append(pkg, O, handle):
    """Build on a first scan then Add_Points(…, false) batches, as radar_odometry.cpp:347,390 does"""
    _, pts, _ = pkg.synth.frame_pair(31, 16, 9000)
    q, _, _ = pkg.synth.frame_pair(32, 512, 16)
    handle.map_build(pts[:3000])
    for s in range(3000, 9000, 2000):
        assert handle.map_add_points(pts[s:s + 2000], False) == 0
    assert handle.map_size() == (9000, 9000)
    _check(handle.map_knn(q, 5, 0.0), O.knn(pts, q, 5, 0.0))
    _check(handle.map_knn(q, 5, 3.0), O.knn(pts, q, 5, 3.0))
