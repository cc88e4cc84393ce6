"""icp4r_set_point_layout: the reference's pcl::PointXYZI rows (32 bytes, intensity at byte 16) and bare xyz rows go in as
they lie in memory, host or device, and every entry point gives the results of the packed x, y, z, w layout bit for bit.
GICP for batched scans == each scan registered alone (and therefore == the oracle, test_gicp_matches_oracle)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def xyzi_rows(p):
    """(n, 4) packed -> (n, 8) rows laid out like pcl::PointXYZI: x y z 1 | intensity 0 0 0"""
    r = np.zeros((p.shape[0], 8), np.float32)
    r[:, :3] = p[:, :3]
    r[:, 3] = 1.0
    r[:, 4] = p[:, 3]
    r[:, 5:] = np.float32(7.5)  # padding garbage must be ignored
    return r


@pytest.mark.parametrize("where", ["host", "device"])
def test_point_layout_equals_packed(pkg, handle, where):
    import torch
    rng = np.random.default_rng(3)
    src, tgt, _ = pkg.synth.frame_pair(21, 900, 1100, extent=25.0)
    src[:, 3] = rng.random(len(src), dtype=np.float32)
    tgt[:, 3] = np.arange(len(tgt), dtype=np.float32)
    put = (lambda a: a) if where == "host" else (lambda a: torch.from_numpy(a).cuda())
    get = (lambda a: a) if where == "host" else (lambda a: a.cpu().numpy() if hasattr(a, "cpu") else a)

    def run(wrap, layout):
        handle.set_point_layout(*layout)
        try:
            out = {}
            handle.map_build(put(wrap(tgt)))
            handle.map_add_points(put(wrap(src[:300])), False)
            idx, d2, found = handle.map_knn(put(wrap(src)), 5, 2.0)
            out["knn"] = (get(idx), get(d2), get(found))
            mp, valid = handle.map_points()
            out["map"] = mp
            o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=5, max_corr_dist=2.0)
            out["reg_map"] = handle.register_map(put(wrap(src)), o)[0]
            o1 = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=8)
            out["reg"] = handle.register(put(wrap(src)), put(wrap(tgt)), o1)[0]
            og = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=4)
            out["gicp"] = handle.register(put(wrap(src)), put(wrap(tgt)), og)[0]
            S, Tg = np.concatenate([src, src[:400]]), np.concatenate([tgt, tgt[:500]])
            so, to = np.array([0, len(src), len(src) + 400], np.int32), np.array([0, len(tgt), len(tgt) + 500], np.int32)
            Tb, _ = handle.register_batch(put(wrap(S)), so if where == "host" else torch.from_numpy(so).cuda(), put(wrap(Tg)),
                                          to if where == "host" else torch.from_numpy(to).cuda(), o1)
            out["batch"] = get(Tb)
            out["vg"] = get(handle.voxel_grid(put(wrap(tgt)), 0.5))
            out["xf"] = get(handle.transform_points(pkg.synth.se3(0.1, 0.0, 0.02, (1.0, 2.0, 0.5)), put(wrap(src))))
            return out
        finally:
            handle.set_point_layout(16, 12)

    ref = run(lambda p: p, (16, 12))
    got = run(xyzi_rows, (32, 16))
    for key in ref:
        a, b = ref[key], got[key]
        if isinstance(a, tuple):
            for x, y in zip(a, b):
                assert np.array_equal(np.asarray(x), np.asarray(y)), key
        elif key in ("reg", "batch"):
            # the resident pair kernel places source points with atomics: poses agree to rounding, not bit for bit, from
            # one run to the next (ICP4R_BATCH_REPRODUCIBLE=1 makes them bit-identical), whatever the input layout
            assert np.abs(np.asarray(a) - np.asarray(b)).max() < 1e-11, key
        else:
            assert np.array_equal(np.asarray(a), np.asarray(b)), key
    # bare x, y, z rows (12 bytes, no w): same geometry, w = 0 in the map
    bare = run(lambda p: np.ascontiguousarray(p[:, :3]), (12, -1))
    assert np.array_equal(bare["knn"][0], ref["knn"][0]) and np.array_equal(bare["knn"][1], ref["knn"][1])
    assert np.array_equal(bare["map"][:, :3], ref["map"][:, :3]) and (bare["map"][:, 3] == 0).all()
    assert np.abs(bare["reg"] - ref["reg"]).max() < 1e-11 and np.array_equal(bare["reg_map"], ref["reg_map"])


def test_point_layout_unaligned_device_rows(pkg, handle):
    """32-byte rows in device memory whose base is only 4-byte aligned (a view into a larger buffer): same answers"""
    import torch
    src, tgt, _ = pkg.synth.frame_pair(22, 700, 900, extent=25.0)
    buf = torch.zeros(8 * len(tgt) + 1, dtype=torch.float32, device="cuda")
    buf[1:] = torch.from_numpy(xyzi_rows(tgt)).reshape(-1).cuda()
    view = buf[1:].reshape(len(tgt), 8)          # data_ptr is base + 4 bytes
    assert view.data_ptr() % 16 != 0 and view.is_contiguous()
    handle.map_build(tgt)
    ref = handle.map_knn(src, 5, 2.0)
    handle.set_point_layout(32, 16)
    try:
        handle.map_build(view)
    finally:
        handle.set_point_layout(16, 12)
    got = handle.map_knn(src, 5, 2.0)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    mp, _ = handle.map_points()
    assert np.array_equal(mp, tgt)


def test_point_layout_rejects_bad_arguments(pkg, handle):
    for stride, woff in ((8, -1), (18, -1), (32, 30), (32, 32), (16, 14), (8192, 0)):
        with pytest.raises(pkg.Icp4rError):
            handle.set_point_layout(stride, woff)
    handle.set_point_layout(16, 12)


def test_gicp_batched_scans_equal_single(pkg, handle):
    """GICP in icp4r_register_map_batch: per-scan covariances, correspondence tables and LM states, one linearisation +
    one LM launch per outer iteration for the whole batch == each scan through icp4r_register_map"""
    import bench
    mp, scans = bench.make_c2()
    scans = [s[: 1200 + 250 * i] for i, s in enumerate(scans[:4])] + [scans[5][:40]]
    off = np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.int32)
    S = np.concatenate(scans)
    handle.map_build(mp)
    T0s = np.stack([pkg.synth.se3(0.002 * i, 0, 0, (0.01 * i, 0, 0)) for i in range(len(scans))])
    for early, iters in ((0, 6), (1, 64)):
        o = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=iters, early_exit=early)
        Tb, rb = handle.register_map_batch(S, off, o, T0s)
        for i, s in enumerate(scans):
            oi = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=iters, early_exit=early, T0=T0s[i])
            Ti, ri, _ = handle.register_map(s, oi)
            assert (rb[i]["converged"], rb[i]["iterations"], rb[i]["n_corr"]) == (ri.converged, ri.iterations, ri.n_corr), i
            assert np.abs(Tb[i] - Ti).max() < 1e-9, (i, np.abs(Tb[i] - Ti).max())
            assert abs(rb[i]["fitness"] - ri.fitness) <= 1e-9 * max(ri.fitness, 1e-12)


def test_brick_normals_equal_warp_per_query_normals(pkg):
    """GICP target covariances: the thread-per-point pass over 3 x 3 x 3 cell blocks (+ the warp-per-query kernel on the
    points it cannot prove) must pick the same neighbours in the same order as the warp-per-query kernel alone — the
    registration that consumes the normals is then bit-identical. Dense map (most points proven) and a sparse cloud
    (most points fall through to the list)."""
    import os
    import bench
    mp, scans = bench.make_c2()
    rng = np.random.default_rng(5)
    sparse = np.zeros((6000, 4), np.float32)
    sparse[:, :3] = rng.uniform(-200, 200, (6000, 3)).astype(np.float32)
    sparse[:40] = sparse[40:80]  # exact duplicates: ties go to the lower index in both kernels
    o = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=6, early_exit=0)
    out = {}
    for mode in ("brick", "warp"):
        if mode == "warp":
            os.environ["ICP4R_NO_BRICK_NORMALS"] = "1"
        try:
            h = pkg.Icp4r(0)
            h.map_build(mp)
            a = h.register_map(scans[0], o)
            h.map_build(sparse)
            b = h.register_map(sparse[:500] + np.float32(0.01), o)
            out[mode] = (a[0], a[1].n_corr, a[1].fitness, b[0], b[1].n_corr, b[1].fitness)
            h.close()
        finally:
            os.environ.pop("ICP4R_NO_BRICK_NORMALS", None)
    for x, y in zip(out["brick"], out["warp"]):
        assert np.array_equal(np.asarray(x), np.asarray(y))
