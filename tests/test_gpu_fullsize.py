"""GPU: BASELINE.json's configurations at their FULL sizes, checked through size-independent properties (the oracle
cannot run these sizes in test time): exhaustive-vs-grid agreement on a sample, sortedness, idempotence, known rigid
offsets recovered, pose chains that must agree between pairs built from the same scene."""
import os
import sys

import numpy as np
from conftest import rot_angle
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_c5_full_size_properties(pkg, O, handle):
    """C5: 16,384-pt scan vs a 20 M-point map"""
    import torch
    import bench
    mp, scans = bench.make_c5(20_000_000)
    d_map = torch.from_numpy(mp).cuda()
    handle.map_build(d_map)
    assert handle.map_size() == (20_000_000, 20_000_000)
    q = scans[0]
    idx, d2, found = handle.map_knn(q, 5, 2.0)
    # ascending distances, inside the gate, indices in range and distinct per row
    f5 = found == 5
    assert f5.mean() > 0.95
    assert (np.diff(d2[f5], axis=1) >= 0).all() and (d2[f5] <= 4.0).all()
    assert (idx[f5] >= 0).all() and (idx[f5] < 20_000_000).all()
    assert all(len(set(r)) == 5 for r in idx[f5][:2000])
    # the reported distances are the float distances to the reported points
    sel = np.nonzero(f5)[0][:4000]
    p = mp[idx[sel]]
    dx = (q[sel, None, :3] - p[:, :, :3]).astype(np.float32)
    want = (dx[..., 0] * dx[..., 0] + dx[..., 1] * dx[..., 1]) + dx[..., 2] * dx[..., 2]
    assert (want.view(np.int32) == d2[sel].view(np.int32)).all()
    # the grid search equals the exhaustive search (same kernel family, no grid) on a sample of queries
    bi, bd, bf = handle.map_knn_brute(q[:128], 5, 2.0)
    assert (bi == idx[:128]).all() and (bd.view(np.int32) == d2[:128].view(np.int32)).all() and (bf == found[:128]).all()
    # ... and the ORACLE's exhaustive exact kNN over all 20 M points (OpenMP over queries, ~5e9 distance evaluations),
    # bit for bit: indices, float distances, counts
    sel = np.concatenate([np.arange(128), np.arange(128, len(q), len(q) // 128)[:128]])
    oi, od, of_ = O.knn(mp, q[sel], 5, 2.0)
    assert (oi == idx[sel]).all() and (od.view(np.int32) == d2[sel].view(np.int32)).all() and (of_ == found[sel]).all()
    # registration: deterministic, and the scan (drawn from the map, moved by a small rigid transform) snaps back
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
    T1, r1, _ = handle.register_map(q, o)
    T2, r2, _ = handle.register_map(q, o)
    assert np.array_equal(T1, T2) and r1.n_corr == r2.n_corr and r1.n_corr > 5000   # planes that fail the 0.2 m flatness test are skipped
    # the fitness the call reports is the mean squared 1-NN distance under the returned pose (PCL getFitnessScore)
    moved = pkg.synth.apply(T1, q)
    i1, dd1, f1 = handle.map_knn(moved, 1, 2.0)
    assert r1.n_fitness == int((f1 == 1).sum())
    assert abs(r1.fitness - float(dd1[f1 == 1, 0].astype(np.float64).mean())) <= 1e-6 * r1.fitness
    # and the loop did not make things worse than the identity pose
    i0, dd0, f0 = handle.map_knn(q, 1, 2.0)
    assert r1.fitness <= float(dd0[f0 == 1, 0].astype(np.float64).mean()) * 1.05
    # one linearisation of the full-size problem against the oracle: the accumulators of a 1,024-point slice of the scan
    # at the returned pose (oracle: exhaustive kNN over the 20 M points), J^T J / J^T r to 1e-9
    oo = O.default_opts(residual=O.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
    part = np.ascontiguousarray(q[:1024])
    got = handle.accumulate_slab(part, o, T1, -1)
    want, _idx, used = O.accumulate(part, mp, oo, T1)
    assert int(got[28]) == used
    assert np.linalg.norm(got[:21] - want[:21]) <= 1e-9 * np.linalg.norm(want[:21])
    g_scale = max(np.linalg.norm(want[21:27]), np.sqrt(want[[0, 6, 11, 15, 18, 20]].sum() * want[27]))
    assert np.linalg.norm(got[21:27] - want[21:27]) <= 1e-9 * g_scale


def test_c4_full_size_properties(pkg, O, handle):
    """C4: 65,536 frame pairs of 2,048 + 2,048 points in one call. Every pair converges by iteration count, matches
    every point (ungated), and a sample of 24 pairs agrees with the ORACLE's registration of the same pair (exhaustive
    exact 1-NN + fp64 Kabsch, 30 iterations) within the stated pose tolerance, and with the same pair registered alone
    through the other implementation of the loop (icp4r_register: grid in global memory, warp per query) to 1e-8."""
    import bench
    src, tgt, so = bench.make_c4(65536)
    o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=30)
    T, res = handle.register_batch(src, so, tgt, so, o)
    assert np.isfinite(T).all() and (res["converged"] == 1).all() and (res["iterations"] == 30).all()
    assert (res["n_corr"] == 2048).all() and (res["n_fitness"] == 2048).all()          # ungated: every point matches
    rng = np.random.default_rng(0)
    oo = O.default_opts(residual=O.P2P_SVD, max_iterations=30)
    single = pkg.Icp4r(0)
    try:
        for p in rng.choice(65536, 24, replace=False):
            a, b = src[so[p]:so[p + 1]], tgt[so[p]:so[p + 1]]
            To, ro, _ = O.register(a, b, oo)
            D = T[p] @ np.linalg.inv(To)
            assert np.linalg.norm(D[:3, 3]) <= 1e-4 and rot_angle(D[:3, :3]) <= 1e-4, (p, np.linalg.norm(D[:3, 3]), rot_angle(D[:3, :3]))
            assert ro.n_corr == res["n_corr"][p] and ro.iterations == res["iterations"][p]
            assert abs(ro.fitness - res["fitness"][p]) <= 1e-6 * max(ro.fitness, 1e-12)
            T1, r1, _ = single.register(a, b, o)
            assert np.abs(T1 - T[p]).max() < 1e-8, (p, np.abs(T1 - T[p]).max())
            assert abs(r1.fitness - res["fitness"][p]) <= 1e-9 * max(r1.fitness, 1e-12) and r1.n_corr == res["n_corr"][p]
            # the fitness is the mean squared 1-NN distance under the returned pose
            moved = pkg.synth.apply(T[p], a)
            single.map_build(b)
            _, dd, ff = single.map_knn(moved, 1, 0.0)
            assert abs(float(dd[:, 0].astype(np.float64).mean()) - res["fitness"][p]) <= 1e-6 * res["fitness"][p]
    finally:
        single.close()


def test_c3_full_length_sequence(pkg, handle):
    """C3: a 2,000-frame odometry sequence through icp4r_odometry_step: the map ends with every scan point in it, in
    order, each scan inserted exactly at its estimated pose; poses are finite; a second run reproduces them bit for bit.
    (How well the synthetic scene constrains the motion is not a property of the implementation and is not asserted.)"""
    seq, gt = pkg.pipeline.synth_sequence(1003, 2000)
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
    poses = pkg.pipeline.run_odometry(handle, seq, o)
    total = sum(len(s) for s in seq)
    assert handle.map_size() == (total, total)
    P = np.array(poses)
    assert np.isfinite(P).all() and len(poses) == 2000
    R = P[:, :3, :3]
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-9 and np.allclose(np.linalg.det(R), 1.0, atol=1e-9)
    pts, valid = handle.map_points()
    assert valid.all()
    offs = np.concatenate([[0], np.cumsum([len(s) for s in seq])])
    for f in (0, 1, 17, 500, 1234, 1999):
        want = handle.transform_points(poses[f], seq[f])
        got = pts[offs[f]:offs[f + 1]]
        assert (got.view(np.int32) == want.view(np.int32)).all(), f
    again = pkg.pipeline.run_odometry(handle, seq, o)
    assert all(np.array_equal(a, b) for a, b in zip(poses, again))
