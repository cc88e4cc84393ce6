"""GPU parity: the scan-to-map node's own per-frame flow (/root/reference/src/radar_odometry.cpp:328,380-429) —
Doppler static-point filter, pointAssociateToMap, Add_Points, Sector_Search, GICP against the sector sub-map, pose
chaining by left multiplication, VoxelGrid over the whole map — through the C ABI with every step on the device,
against the same flow on the CPU: the reference's ikd-Tree (compiled unmodified, when it travelled with the repo) for
Add_Points / Sector_Search and the oracle's restated filter, GICP loop and voxel grid."""
import numpy as np
import pytest
from conftest import rot_angle

pytestmark = pytest.mark.gpu


def oracle_flow(O, pkg, frames, oo, priors, radius=80.0, leaf=0.5, seed=1):
    odom = np.eye(4)
    poses, subs = [], []
    total = sum(len(f) for f in frames)
    buf = np.empty((total, 4), np.float32)
    n = 0
    tree = O.IkdTree() if O.have_ref() else None
    om = O.OracleMap(total) if tree is None else None
    for f, rec in enumerate(frames):
        T = np.asarray(priors[f], np.float64)
        mask, _ = O.doppler_filter(rec, 0, seed + f)
        static = np.ascontiguousarray(rec[mask.astype(bool)][:, :4])
        scan_w, _ = O.transform(T, static)
        buf[n:n + len(scan_w)] = scan_w
        if om is not None:
            om.add_points(scan_w, False)
        if f == 0:
            if tree is not None:
                tree.build(scan_w)
            n += len(scan_w)
            odom = T.copy()
        else:
            if tree is not None:
                tree.add_points(scan_w, False)
            n += len(scan_w)
            yaw = pkg.pipeline.yaw_deg(T)
            if tree is not None:
                idx = np.sort(tree.sector(T[:3, 3], radius, yaw))
            else:
                idx = np.sort(om.sector(T[:3, 3], radius, yaw))
            subs.append(idx)
            D, _r, _ = O.gicp_register(scan_w, np.ascontiguousarray(buf[idx]), oo)
            odom = D @ T
        poses.append(odom.copy())
    ds = O.voxel_grid(buf[:n], leaf)
    if tree is not None:
        tree.close()
    return poses, subs, ds


def test_reference_flow_matches_cpu(pkg, O, handle):
    import torch
    frames, gt = pkg.pipeline.synth_radar_sequence(31, 6, pts_per_frame=1500, fov_deg=55.0, max_range=78.0, scan_radius=90.0, forward="y")
    # the node places every scan with a prior pose (odometry / ground truth queue); perturb it so that GICP has work to do
    rng = np.random.default_rng(5)
    gt = [g @ pkg.synth.random_small_se3(rng, 0.05, 0.3) for g in gt]
    o = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=8, early_exit=0, max_corr_dist=0.0)
    oo = O.default_opts(residual=O.GICP, k=5, max_iterations=8, early_exit=0, max_corr_dist=0.0)
    want, subs, ds_want = oracle_flow(O, pkg, frames, oo, gt)
    # step by step on the device with the sector indices sorted (the sub-map order is the only thing the two sides may
    # legitimately differ in: Sector_Search returns tree order, the device an unspecified order)
    dev = torch.device("cuda", 0)
    odom = np.eye(4)
    for f, rec in enumerate(frames):
        T = gt[f]
        static, _ = handle.doppler_static_points(torch.from_numpy(rec).to(dev), 0, seed=1 + f)
        scan_w = handle.transform_points(T, static)
        if f == 0:
            handle.map_build(scan_w)
            odom = T.copy()
        else:
            handle.map_add_points(scan_w, False)
            idx = torch.sort(handle.map_sector_dev(T[:3, 3], 80.0, pkg.pipeline.yaw_deg(T)))[0].contiguous()
            assert np.array_equal(idx.cpu().numpy(), subs[f - 1]), f          # the same sub-map, as a set
            D, res = handle.register_submap(scan_w, idx, o)
            odom = D @ T
        E = odom @ np.linalg.inv(want[f])
        assert np.linalg.norm(E[:3, 3]) <= 1e-4 and rot_angle(E[:3, :3]) <= 1e-4, (f, np.linalg.norm(E[:3, 3]), rot_angle(E[:3, :3]))
    ds = handle.voxel_grid(None, 0.5)
    assert ds.shape == ds_want.shape and np.abs(ds - ds_want).max() <= 1e-4
    # the packaged loop (device-resident indices in whatever order the sector search produced them) gives the same poses
    poses, n_ds = pkg.pipeline.run_reference_flow(handle, [torch.from_numpy(r).to(dev) for r in frames], o, priors=gt)
    for f in range(len(frames)):
        E = poses[f] @ np.linalg.inv(want[f])
        assert np.linalg.norm(E[:3, 3]) <= 1e-4 and rot_angle(E[:3, :3]) <= 1e-4, f
    assert n_ds == len(ds_want)
    # host buffers through the same calls
    poses_h, _ = pkg.pipeline.run_reference_flow(handle, frames, o, priors=gt)
    for f in range(len(frames)):
        assert np.abs(poses_h[f] - poses[f]).max() <= 1e-6


def test_static_points_and_submap_edges(pkg, O, handle):
    rec = pkg.pipeline.synth_radar_sequence(5, 1, pts_per_frame=700)[0][0]
    mask, _ = handle.doppler_filter(rec, 0, seed=3)
    static, res = handle.doppler_static_points(rec, 0, seed=3)
    assert np.array_equal(static, rec[np.asarray(mask).astype(bool)][:, :4]) and res.n_static == len(static)
    empty, _ = handle.doppler_static_points(rec[:0], 0, seed=3)
    assert empty.shape == (0, 4)
    # sub-map registration == registration against the same points passed explicitly; out-of-range indices are ignored
    src, tgt, _ = pkg.synth.frame_pair(12, 600, 3000, extent=30.0)
    handle.map_build(tgt)
    idx = np.arange(0, 3000, 2, dtype=np.int32)
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=6, max_corr_dist=2.0)
    T1, r1 = handle.register_submap(src, idx, o)
    T2, r2, _ = handle.register(src, tgt[idx], o)
    assert np.array_equal(T1, T2) and r1.n_corr == r2.n_corr
    bad = np.concatenate([idx, np.array([-5, 10 ** 7], np.int32)])
    T3, r3 = handle.register_submap(src, bad, o)
    assert np.abs(T3 - T1).max() <= 1e-12 and r3.n_corr == r1.n_corr
    with pytest.raises(pkg.api.Icp4rError):
        pkg.Icp4r(0).register_submap(src, idx, o)   # no map on that handle
