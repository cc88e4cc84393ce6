"""GPU parity: registration loops vs the CPU oracle, through the C ABI.

Bars (north_star): correspondence indices bit-exact; per-iteration J^T J / J^T r relative error <= 1e-5;
final pose within 1e-4 m / 1e-4 rad after the fixed iteration count.
"""
import numpy as np
from conftest import rot_angle
import pytest

pytestmark = pytest.mark.gpu

POSE_TOL_T = 1e-4   # metres
POSE_TOL_R = 1e-4   # radians
ACC_RTOL = 1e-5


def pose_err(A, B):
    D = A @ np.linalg.inv(B)
    ang = rot_angle(D[:3, :3])
    return np.linalg.norm(D[:3, 3]), ang


def check_dumps(O, kind, k, src, tgt, oo, bufs, iterations):
    """For each iteration the library ran: recompute correspondences and accumulators on the CPU from the
    pose the GPU reports for that iteration. Indices must match exactly, accumulators to 1e-5."""
    dp, da, di = bufs
    for it in range(iterations):
        acc, idx, used = O.accumulate(src, tgt, oo, dp[it].reshape(4, 4))
        assert (idx == di[it]).all(), f"iteration {it}: {np.count_nonzero((idx != di[it]).any(axis=1))} correspondences differ"
        nv = 17 if kind == O.P2P_SVD else 29
        a, b = da[it][:nv], acc[:nv]
        if kind == O.P2P_SVD:
            assert np.linalg.norm(a - b) <= ACC_RTOL * np.linalg.norm(b)
        else:
            assert np.linalg.norm(a[:21] - b[:21]) <= ACC_RTOL * np.linalg.norm(b[:21]), f"JtJ it {it}"
            # J^T r vanishes at the optimum (a sum of cancelling terms): measure its error against the natural scale of
            # those terms, |J^T r| <= sqrt(tr(J^T J) * sum r^2), not against the vanishing net value
            diag = [0, 6, 11, 15, 18, 20]
            g_scale = max(np.linalg.norm(b[21:27]), np.sqrt(b[diag].sum() * b[27]))
            assert np.linalg.norm(a[21:27] - b[21:27]) <= ACC_RTOL * max(g_scale, 1e-12), f"Jtr it {it}"
            assert a[28] == b[28]


KINDS = [("P2P_SVD", 1, 0.0), ("P2P_GN", 1, 0.0), ("P2PLANE_KNN", 5, 2.0), ("P2LINE", 2, 3.0), ("P2P_SVD", 1, 2.5),
         ("P2PLANE_3PT", 3, 2.0), ("P2PLANE_KNN", 8, 3.0)]


@pytest.mark.parametrize("name,k,gate", KINDS)
def test_register_pair_matches_oracle(pkg, O, handle, name, k, gate):
    kind = getattr(pkg, name)
    src, tgt, _ = pkg.synth.frame_pair(1001, 1024, 4000, extent=40.0)
    iters = 12
    o = pkg.default_opts(residual=kind, k=k, max_iterations=iters, max_corr_dist=gate)
    oo = O.default_opts(residual=kind, k=k, max_iterations=iters, max_corr_dist=gate)
    T, res, bufs = handle.register(src, tgt, o, dump=True)
    To, ro, _ = O.register(src, tgt, oo)
    assert res.iterations == ro.iterations == iters and res.converged == ro.converged
    check_dumps(O, kind, k, src, tgt, oo, bufs, res.iterations)
    et, er = pose_err(T, To)
    assert et <= POSE_TOL_T and er <= POSE_TOL_R, (et, er)
    assert res.n_corr == ro.n_corr and res.n_fitness == ro.n_fitness
    assert abs(res.fitness - ro.fitness) <= 1e-6 * max(ro.fitness, 1e-12)


@pytest.mark.parametrize("name,k,gate,s", [("P2LINE", 2, 3.0, 0.5), ("P2PLANE_3PT", 3, 2.0, 0.5), ("P2PLANE_3PT", 3, 2.5, 0.25), ("P2LINE", 2, 3.0, 1.0)])
def test_interpolated_functors_match_oracle(pkg, O, handle, name, k, gate, s):
    """RadarEdgeFactor / LidarPlaneFactor with their interpolation ratio s (radarFactor.hpp:26-32,78-84): the point is placed
    with slerp(I, q, s) p + s t. Device: closed-form Jacobian (left Jacobians of SO(3)); oracle: the functor restated
    literally and differentiated numerically. Correspondences bit-exact, J^T J / J^T r to 1e-5, final pose to 1e-4."""
    kind = getattr(pkg, name)
    src, tgt, _ = pkg.synth.frame_pair(1001, 1024, 4000, extent=40.0)
    T0 = pkg.synth.se3(0.02, 0.004, -0.003, (0.1, -0.05, 0.02))
    iters = 8
    o = pkg.default_opts(residual=kind, k=k, max_iterations=iters, max_corr_dist=gate, interp_s=s, T0=T0)
    oo = O.default_opts(residual=kind, k=k, max_iterations=iters, max_corr_dist=gate, interp_s=s, T0=T0)
    T, res, bufs = handle.register(src, tgt, o, dump=True)
    To, ro, _ = O.register(src, tgt, oo)
    assert res.iterations == ro.iterations == iters
    check_dumps(O, kind, k, src, tgt, oo, bufs, res.iterations)
    et, er = pose_err(T, To)
    assert et <= POSE_TOL_T and er <= POSE_TOL_R, (et, er)
    assert res.n_corr == ro.n_corr
    # the map-resident and batched-scan entry points take the same option
    handle.map_build(tgt)
    T2, r2, _ = handle.register_map(src, o)
    assert np.abs(T2 - T).max() <= 1e-9 and r2.n_corr == res.n_corr
    if s != 1.0:
        o1 = pkg.default_opts(residual=kind, k=k, max_iterations=iters, max_corr_dist=gate, T0=T0)
        T1, _r1, _ = handle.register_map(src, o1)
        assert np.abs(T1 - T).max() > 1e-4   # s really changes the problem


def test_c1_config(pkg, O, handle):
    """BASELINE config 1: 1,024-pt frame pair, point-to-point, 30 iterations, ungated"""
    src, tgt, _ = pkg.synth.frame_pair(1001, 1024)
    o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=30)
    oo = O.default_opts(residual=O.P2P_SVD, max_iterations=30)
    T, res, bufs = handle.register(src, tgt, o, dump=True)
    To, ro, _ = O.register(src, tgt, oo)
    check_dumps(O, O.P2P_SVD, 1, src, tgt, oo, bufs, 30)
    et, er = pose_err(T, To)
    assert et <= POSE_TOL_T and er <= POSE_TOL_R, (et, er)
    # the batched (shared-memory resident) kernel must agree with the map kernel and the oracle
    off = np.array([0, 1024], np.int32)
    Tb, rb = handle.register_batch(src, off, tgt, off, o)
    et, er = pose_err(Tb[0], To)
    assert et <= POSE_TOL_T and er <= POSE_TOL_R, (et, er)
    assert rb["iterations"][0] == 30 and rb["converged"][0] == 1 and rb["n_corr"][0] == ro.n_corr
    assert abs(rb["fitness"][0] - ro.fitness) <= 1e-6 * ro.fitness


def test_c2_config(pkg, O, handle):
    """BASELINE config 2: 4,096-pt scan vs 200k-pt map, k=5 point-to-plane, 20 iterations, gate 2 m.
    Oracle kNN comes from the reference's own ikd-Tree when the compiled reference travelled with the repo."""
    scan, mp, _ = pkg.synth.scan_to_map(1002, 4096, 200000)
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
    oo = O.default_opts(residual=O.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
    handle.map_build(mp)
    T, res, bufs = handle.register_map(scan, o, dump=True)
    searcher = None
    if O.have_ref():
        searcher = O.IkdTree(nthreads=8)
        searcher.build(mp)
    To, ro, _ = O.register(scan, mp, oo, searcher=searcher)
    et, er = pose_err(T, To)
    assert et <= POSE_TOL_T and er <= POSE_TOL_R, (et, er)
    assert res.n_corr == ro.n_corr
    dp, da, di = bufs
    for it in (0, 1, 7, 19):
        acc, idx, used = O.accumulate(scan, mp, oo, dp[it].reshape(4, 4), searcher=searcher)
        assert (idx == di[it]).all()
        assert np.linalg.norm(da[it][:21] - acc[:21]) <= ACC_RTOL * np.linalg.norm(acc[:21])
        assert np.linalg.norm(da[it][21:27] - acc[21:27]) <= ACC_RTOL * np.linalg.norm(acc[21:27])


def test_early_exit_and_degenerate(pkg, O, handle):
    src, tgt, _ = pkg.synth.frame_pair(77, 800, 3000, extent=30.0)
    for kind, k, gate in ((pkg.P2P_SVD, 1, 0.0), (pkg.P2PLANE_KNN, 5, 2.0)):
        o = pkg.default_opts(residual=kind, k=k, max_iterations=60, early_exit=1, max_corr_dist=gate, mse_abs_eps=1e-9)
        oo = O.default_opts(residual=kind, k=k, max_iterations=60, early_exit=1, max_corr_dist=gate, mse_abs_eps=1e-9)
        T, res, _ = handle.register(src, tgt, o)
        To, ro, _ = O.register(src, tgt, oo)
        assert (res.converged, res.iterations) == (ro.converged, ro.iterations)
        if kind == pkg.P2P_SVD:
            assert res.iterations < 60
        et, er = pose_err(T, To)
        assert et <= POSE_TOL_T and er <= POSE_TOL_R
    # no correspondences at all: gate far smaller than the offset -> not converged, pose = initial guess
    far = src.copy()
    far[:, :3] += 1000.0
    o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=5, max_corr_dist=0.5)
    T, res, _ = handle.register(far, tgt, o)
    assert res.converged == 0 and res.iterations == 0 and res.n_corr == 0 and np.allclose(T, np.eye(4))
    assert np.isinf(res.fitness) and res.n_fitness == 0
    # empty source / empty target
    T, res, _ = handle.register(src[:0], tgt, pkg.default_opts(max_iterations=3))
    assert res.converged == 0 and np.allclose(T, np.eye(4))
    T, res, _ = handle.register(src, tgt[:0], pkg.default_opts(max_iterations=3))
    assert res.converged == 0 and np.allclose(T, np.eye(4))
    # initial guess is honoured
    T0 = pkg.synth.se3(0.02, 0, 0, (0.3, -0.2, 0.0))
    o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=4, T0=T0)
    oo = O.default_opts(residual=O.P2P_SVD, max_iterations=4, T0=T0)
    T, res, _ = handle.register(src, tgt, o)
    To, ro, _ = O.register(src, tgt, oo)
    et, er = pose_err(T, To)
    assert et <= POSE_TOL_T and er <= POSE_TOL_R


def test_batch_ragged(pkg, O, handle):
    """batched registration over pairs of different sizes, both residual kinds, gated and ungated"""
    rng = np.random.default_rng(9)
    sizes = [(300, 500), (1024, 1024), (17, 2048), (2048, 33), (700, 700), (1, 5), (64, 0), (0, 64)]
    srcs, tgts = [], []
    for i, (n, m) in enumerate(sizes):
        s, t, _ = pkg.synth.frame_pair(100 + i, max(n, 1), max(m, 1), extent=float(rng.choice([20.0, 80.0])))
        srcs.append(s[:n])
        tgts.append(t[:m])
    soff = np.concatenate([[0], np.cumsum([len(s) for s in srcs])]).astype(np.int32)
    toff = np.concatenate([[0], np.cumsum([len(t) for t in tgts])]).astype(np.int32)
    S, Tg = np.concatenate(srcs), np.concatenate(tgts)
    for kind, gate in ((pkg.P2P_SVD, 0.0), (pkg.P2P_GN, 0.0), (pkg.P2P_SVD, 4.0)):
        o = pkg.default_opts(residual=kind, max_iterations=8, max_corr_dist=gate)
        oo = O.default_opts(residual=kind, max_iterations=8, max_corr_dist=gate)
        Tb, rb = handle.register_batch(S, soff, Tg, toff, o)
        for i in range(len(sizes)):
            To, ro, _ = O.register(srcs[i], tgts[i], oo)
            assert (rb["converged"][i], rb["iterations"][i], rb["n_corr"][i]) == (ro.converged, ro.iterations, ro.n_corr), i
            et, er = pose_err(Tb[i], To)
            assert et <= POSE_TOL_T and er <= POSE_TOL_R, (i, et, er)
            if ro.n_fitness:
                assert abs(rb["fitness"][i] - ro.fitness) <= 1e-6 * ro.fitness


def test_batch_pairs_larger_than_shared_memory(pkg, O, handle):
    """pairs above one SM's shared memory (here 9,000 + 9,000 points) take the map path pair by pair instead of failing"""
    import torch
    srcs, tgts = [], []
    for i in range(3):
        s, t, _ = pkg.synth.frame_pair(300 + i, 9000 if i == 1 else 500, 9000 if i == 1 else 700, extent=60.0)
        srcs.append(s)
        tgts.append(t)
    soff = np.concatenate([[0], np.cumsum([len(s) for s in srcs])]).astype(np.int32)
    toff = np.concatenate([[0], np.cumsum([len(t) for t in tgts])]).astype(np.int32)
    S, Tg = np.concatenate(srcs), np.concatenate(tgts)
    o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=6)
    oo = O.default_opts(residual=O.P2P_SVD, max_iterations=6)
    Tb, rb = handle.register_batch(S, soff, Tg, toff, o)
    Td, rd = handle.register_batch(torch.from_numpy(S).cuda(), torch.from_numpy(soff).cuda(), torch.from_numpy(Tg).cuda(),
                                   torch.from_numpy(toff).cuda(), o)
    assert np.array_equal(Td.cpu().numpy().reshape(-1, 4, 4), Tb)
    for i in range(3):
        To, ro, _ = O.register(srcs[i], tgts[i], oo)
        et, er = pose_err(Tb[i], To)
        assert et <= POSE_TOL_T and er <= POSE_TOL_R, (i, et, er)
        assert rb["n_corr"][i] == ro.n_corr and rb["iterations"][i] == ro.iterations


def test_transform_points(pkg, O, handle):
    src, _, Tgt = pkg.synth.frame_pair(5, 999)
    got = handle.transform_points(Tgt, src)
    want, _ = O.transform(Tgt, src)
    assert (got.view(np.int32) == want.view(np.int32)).all()


def test_gicp_matches_oracle(pkg, O, handle):
    """fast_gicp cost (what radar_odometry.cpp:399-405 runs, k = 5): per-iteration H / g / cost at the GPU's own poses
    against the oracle's independent formulation (explicit Mahalanobis inverse), correspondences bit-exact, final
    pose within tolerance, same Levenberg-Marquardt trajectory (iterations, convergence flag)."""
    src, tgt, _ = pkg.synth.frame_pair(5, 1500, 6000, extent=25.0)
    sn, tn = O.gicp_normals(src, 5), O.gicp_normals(tgt, 5)
    for early, iters in ((0, 8), (1, 64)):
        o = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=iters, early_exit=early)
        oo = O.default_opts(residual=O.GICP, k=5, max_iterations=iters, early_exit=early)
        T, res, bufs = handle.register(src, tgt, o, dump=True)
        To, ro, _ = O.gicp_register(src, tgt, oo, normals=(sn, tn))
        assert (res.converged, res.iterations, res.n_corr) == (ro.converged, ro.iterations, ro.n_corr)
        dp, da, di = bufs
        for it in range(res.iterations):
            acc, idx, used = O.gicp_linearize(src, sn, tgt, tn, oo, dp[it].reshape(4, 4))
            assert (idx == di[it][:, 0]).all(), it
            assert np.linalg.norm(da[it][:21] - acc[:21]) <= ACC_RTOL * np.linalg.norm(acc[:21]), it
            assert np.linalg.norm(da[it][21:27] - acc[21:27]) <= ACC_RTOL * max(np.linalg.norm(acc[21:27]), 1e-9), it
            assert abs(da[it][27] - acc[27]) <= ACC_RTOL * acc[27] and da[it][28] == acc[28]
        et, er = pose_err(T, To)
        assert et <= POSE_TOL_T and er <= POSE_TOL_R, (et, er)
        assert abs(res.fitness - ro.fitness) <= 1e-6 * ro.fitness
    # against the resident map as well (the scan-to-map call shape), normals cached on the map
    handle.map_build(tgt)
    o = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=64, early_exit=1)
    T2, r2, _ = handle.register_map(src, o)
    T3, r3, _ = handle.register_map(src, o)
    assert np.array_equal(T2, T3) and np.abs(T2 - T).max() < 1e-12


def test_register_map_batch(pkg, O, handle):
    """several scans against the same map in one launch sequence == each scan registered alone"""
    import bench
    mp, scans = bench.make_c2()
    scans = [s[: 1500 + 300 * i] for i, s in enumerate(scans[:5])] + [scans[5][:0], scans[6][:7]]   # ragged, empty, tiny
    off = np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.int32)
    S = np.concatenate(scans)
    handle.map_build(mp)
    for kind, k, gate in ((pkg.P2PLANE_KNN, 5, 2.0), (pkg.P2P_SVD, 1, 2.0)):
        o = pkg.default_opts(residual=kind, k=k, max_iterations=6, max_corr_dist=gate)
        T0s = np.stack([pkg.synth.se3(0.002 * i, 0, 0, (0.01 * i, 0, 0)) for i in range(len(scans))])
        Tb, rb = handle.register_map_batch(S, off, o, T0s)
        for i, s in enumerate(scans):
            oi = pkg.default_opts(residual=kind, k=k, max_iterations=6, max_corr_dist=gate, T0=T0s[i])
            Ti, ri, _ = handle.register_map(s, oi)
            assert np.abs(Tb[i] - Ti).max() < 1e-9, (i, np.abs(Tb[i] - Ti).max())
            assert (rb["converged"][i], rb["iterations"][i], rb["n_corr"][i], rb["n_fitness"][i]) == (ri.converged, ri.iterations, ri.n_corr, ri.n_fitness)
        # one of them against the oracle too
        oo = O.default_opts(residual=kind, k=k, max_iterations=6, max_corr_dist=gate, T0=T0s[2])
        To, ro, _ = O.register(scans[2], mp, oo)
        et, er = pose_err(Tb[2], To)
        assert et <= POSE_TOL_T and er <= POSE_TOL_R
    # device-resident scans
    import torch
    Tb2, rb2 = handle.register_map_batch(torch.from_numpy(S).cuda(), off, o, T0s)
    assert np.array_equal(Tb2, Tb)


def test_non_finite_points_are_ignored(pkg, O, handle):
    """NaN / inf coordinates in the scan or in the map never match anything: the result equals the registration of the
    clouds with those points removed (indices keep referring to the original arrays)"""
    src, tgt, _ = pkg.synth.frame_pair(91, 900, 4000, extent=30.0)
    src_bad, tgt_bad = src.copy(), tgt.copy()
    src_bad[::37, 0] = np.nan
    src_bad[5::53, 2] = np.inf
    tgt_bad[::29, 1] = np.nan
    tgt_bad[7::61, 0] = -np.inf
    keep_s = np.isfinite(src_bad[:, :3]).all(1)
    keep_t = np.isfinite(tgt_bad[:, :3]).all(1)
    for kind, k in ((pkg.P2P_SVD, 1), (pkg.P2PLANE_KNN, 5)):
        o = pkg.default_opts(residual=kind, k=k, max_iterations=8, max_corr_dist=2.0)
        T_bad, r_bad, _ = handle.register(src_bad, tgt_bad, o)
        T_ok, r_ok, _ = handle.register(src_bad[keep_s], tgt_bad[keep_t], o)
        assert np.isfinite(T_bad).all()
        assert r_bad.n_corr == r_ok.n_corr and r_bad.n_fitness == r_ok.n_fitness
        et, er = pose_err(T_bad, T_ok)
        assert et <= 1e-9 and er <= 1e-9
        handle.map_build(tgt_bad)
        T_map, r_map, _ = handle.register_map(src_bad, o)
        et, er = pose_err(T_map, T_ok)
        assert et <= 1e-9 and er <= 1e-9 and r_map.n_corr == r_ok.n_corr


def test_large_host_batch_is_pipelined_and_equal(pkg, monkeypatch):
    """host-resident batches above 32 MB are copied in chunks on a second stream while earlier chunks are registered;
    the poses must equal the device-resident (single launch) path — bit for bit with ICP4R_BATCH_REPRODUCIBLE=1 (the
    source cloud is placed into its cells in index order), to summation-order rounding otherwise — ragged pair sizes
    included"""
    import torch
    monkeypatch.setenv("ICP4R_BATCH_REPRODUCIBLE", "1")
    handle = pkg.Icp4r(0)
    monkeypatch.delenv("ICP4R_BATCH_REPRODUCIBLE")
    fast = pkg.Icp4r(0)
    rng = np.random.default_rng(5)
    base = [pkg.synth.frame_pair(300 + i, 2048) for i in range(8)]
    n_pairs = 700
    srcs, tgts = [], []
    for p in range(n_pairs):
        s, t, _ = base[p % 8]
        cut_s, cut_t = 2048 - int(rng.integers(0, 64)), 2048 - int(rng.integers(0, 64))
        srcs.append(s[:cut_s])
        tgts.append(t[:cut_t])
    so = np.concatenate([[0], np.cumsum([len(s) for s in srcs])]).astype(np.int32)
    to = np.concatenate([[0], np.cumsum([len(t) for t in tgts])]).astype(np.int32)
    S, T = np.concatenate(srcs), np.concatenate(tgts)
    assert (S.nbytes + T.nbytes) > (32 << 20)
    o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=6)
    Th, rh = handle.register_batch(S, so, T, to, o)
    Td, rd = handle.register_batch(torch.from_numpy(S).cuda(), torch.from_numpy(so).cuda(), torch.from_numpy(T).cuda(), torch.from_numpy(to).cuda(), o)
    handle.synchronize()  # device-resident calls are asynchronous on the handle's stream
    assert np.array_equal(Th.reshape(n_pairs, 16), Td.cpu().numpy())
    Th2, _ = handle.register_batch(S, so, T, to, o)
    assert np.array_equal(Th, Th2)
    Tf, _ = fast.register_batch(S, so, T, to, o)          # default placement: same poses up to summation order
    assert np.allclose(Tf, Th, rtol=0, atol=1e-10)
    handle.close()
    fast.close()
    rdn = rd.cpu().numpy().view(pkg.api.RESULT_DTYPE).reshape(-1)
    assert (rh["n_corr"] == rdn["n_corr"]).all() and (rh["iterations"] == rdn["iterations"]).all()


def test_persistent_loop_equals_per_iteration_launches(pkg):
    """ICP4R_PERSIST=1: single-scan loops run as ONE cooperative launch (reg_loop_kernel: grid-wide hand-over between
    iterations inside the kernel); the result must be bit-identical to the default one-launch-per-iteration graph, with
    and without early exit, for every residual kind of the scan-to-map call"""
    import os
    import bench
    mp, scans = bench.make_c2()
    scan = scans[0][:3000]
    out = {}
    for mode in ("persist", "launches"):
        if mode == "persist":
            os.environ["ICP4R_PERSIST"] = "1"
        try:
            h = pkg.Icp4r(0)
        finally:
            os.environ.pop("ICP4R_PERSIST", None)
        h.map_build(mp)
        rows = []
        for kind, k in ((pkg.P2PLANE_KNN, 5), (pkg.P2PLANE_KNN, 8), (pkg.P2P_SVD, 1), (pkg.P2P_GN, 1), (pkg.P2LINE, 2), (pkg.P2PLANE_3PT, 3)):
            for early, iters in ((0, 12), (1, 40)):
                o = pkg.default_opts(residual=kind, k=k, max_iterations=iters, early_exit=early, max_corr_dist=2.0,
                                     rot_eps=1e-5, trans_eps=1e-5, mse_abs_eps=1e-9)
                n0 = h.launch_count()
                T, r, _ = h.register_map(scan, o)
                T2, r2, _ = h.register_map(scan, o)   # and again: the hand-over word starts from zero every call
                assert np.array_equal(T, T2)
                rows.append((T, r.converged, r.iterations, r.n_corr, r.fitness, h.launch_count() - n0))
        out[mode] = rows
        h.close()
    for a, b in zip(out["persist"], out["launches"]):
        assert np.array_equal(a[0], b[0]) and a[1:5] == b[1:5], (a, b)
        assert a[5] == 2 * 3, a[5]          # init + loop + fitness, twice
        assert b[5] > a[5]
