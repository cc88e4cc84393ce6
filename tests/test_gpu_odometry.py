"""GPU parity: the scan-to-map odometry loop (BASELINE config 3, scaled down) — register against the growing map,
transform, Add_Points(false) — against the same loop on the CPU oracle with the reference's ikd-Tree as the map."""
import numpy as np
from conftest import rot_angle
import pytest

pytestmark = pytest.mark.gpu


def oracle_odometry(O, scans, oo):
    T = np.eye(4)
    poses = [T.copy()]
    w, _ = O.transform(T, scans[0])
    use_ref = O.have_ref()
    if use_ref:
        tree = O.IkdTree(nthreads=4)   # KD_TREE<PointXYZI>(0.3, 0.6, 0.5), radar_odometry.cpp:92
        tree.build(w)
    pts = [w]
    for scan in scans[1:]:
        for i in range(16):
            oo.T0[i] = float(T.reshape(16)[i])
        mp = np.concatenate(pts)
        T, res, _ = O.register(scan, mp, oo, searcher=tree if use_ref else None)
        poses.append(T.copy())
        w, _ = O.transform(T, scan)
        pts.append(w)
        if use_ref:
            assert tree.add_points(w, False) == 0
    if use_ref:
        tree.close()
    return poses, sum(len(p) for p in pts)


def test_odometry_sequence_matches_oracle(pkg, O, handle):
    scans, gt = pkg.pipeline.synth_sequence(1003, 14, pts_per_scan=900, raw_per_scan=1200, extent=120.0, scan_radius=40.0)
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=10, max_corr_dist=2.0)
    oo = O.default_opts(residual=O.P2PLANE_KNN, k=5, max_iterations=10, max_corr_dist=2.0)
    got = pkg.pipeline.run_odometry(handle, scans, o)
    want, total = oracle_odometry(O, scans, oo)
    assert handle.map_size() == (total, total)
    for f, (a, b) in enumerate(zip(got, want)):
        D = a @ np.linalg.inv(b)
        ang = rot_angle(D[:3, :3])
        assert np.linalg.norm(D[:3, 3]) <= 1e-4 and ang <= 1e-4, (f, np.linalg.norm(D[:3, 3]), ang)
    # and the loop really tracks the trajectory (sanity of the synthetic sequence, not a parity bar)
    drift = np.linalg.norm(got[-1][:3, 3] - (np.linalg.inv(gt[0]) @ gt[-1])[:3, 3])
    assert drift < 3.0, drift  # ~5.6 m travelled; the ground plane constrains x, y only through sparse walls


def test_fused_odometry_step_equals_separate_calls(pkg, handle):
    """icp4r_odometry_step (register + transform + Add_Points in one call, device-side) gives the poses and the map of
    the three separate C-ABI calls, bit for bit; host and device inputs alike"""
    import torch
    scans, _ = pkg.pipeline.synth_sequence(77, 10, pts_per_scan=800, raw_per_scan=1000, extent=100.0, scan_radius=40.0)
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=8, max_corr_dist=2.0)
    sep = pkg.pipeline.run_odometry(handle, scans, o, fused=False)
    pts_sep, valid_sep = handle.map_points()
    fused = pkg.pipeline.run_odometry(handle, scans, o, fused=True)
    pts_fused, valid_fused = handle.map_points()
    for a, b in zip(sep, fused):
        assert np.array_equal(a, b)
    assert np.array_equal(pts_sep.view(np.int32), pts_fused.view(np.int32)) and np.array_equal(valid_sep, valid_fused)
    dev = pkg.pipeline.run_odometry(handle, [torch.from_numpy(s).cuda() for s in scans], o, fused=True)
    for a, b in zip(sep, dev):
        assert np.array_equal(a, b)


def test_scan_to_scan_node_matches_oracle(pkg, O, handle):
    """the scan-to-scan node (Doppler filter -> ICP against the previous frame -> right-multiplied pose chain) over
    the C ABI, batched replay and frame by frame, against the oracle running the same steps"""
    frames, gt = pkg.pipeline.synth_radar_sequence(31, 9, pts_per_frame=900)
    o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=10)            # PCL defaults of the node (ungated, 10 iterations)
    oo = O.default_opts(residual=O.P2P_SVD, max_iterations=10)
    poses_b, vel_b, res_b = pkg.pipeline.run_scan_to_scan(handle, frames, o, batched=True, seed=5)
    poses_s, vel_s, res_s = pkg.pipeline.run_scan_to_scan(handle, frames, o, batched=False, seed=5)
    # oracle chain
    statics, vel_o = [], []
    for f, rec in enumerate(frames):
        m, out = O.doppler_filter(rec, 0, seed=5 + f)
        statics.append(np.ascontiguousarray(rec[m.astype(bool)][:, :4]))
        vel_o.append(np.array(list(out.v)))
    cur, poses_o = np.eye(4), [np.eye(4)]
    for f in range(1, len(frames)):
        T, r, _ = O.register(statics[f], statics[f - 1], oo)
        cur = cur @ T
        poses_o.append(cur.copy())
    for f in range(len(frames)):
        assert np.allclose(vel_b[f], vel_o[f], rtol=1e-9, atol=1e-12)
        for got in (poses_b[f], poses_s[f]):
            D = got @ np.linalg.inv(poses_o[f])
            ang = rot_angle(D[:3, :3])
            assert np.linalg.norm(D[:3, 3]) <= 1e-4 and ang <= 1e-4, f
    # the least-squares velocity is that of the static world relative to the sensor (the reference solves K v = v_r
    # with v_r = -u . v_ego): opposite to the motion along +x, ~0.4 m per 0.1 s frame
    assert -5.5 < vel_b[4][0] < -2.5
