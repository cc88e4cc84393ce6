"""Scan-to-map odometry loop over the C ABI (BASELINE config 3): what /root/reference/src/radar_odometry.cpp:344-421
does per radar frame, with the map on the device.

  first frame : KD_TREE::Build                      -> icp4r_map_build
  every frame : align(scan -> map, guess = last pose) -> icp4r_register_map   (the reference's FastGICP call)
                pointAssociateToMap                  -> icp4r_transform_points
                Add_Points(scan_world, false)        -> icp4r_map_add_points

The reference inserts the scan with the ground-truth pose BEFORE aligning it (radar_odometry.cpp:390 then :399), so
its target already contains the source; here the scan is aligned first and inserted with the ESTIMATED pose, which
is what an odometry without ground truth has to do.
"""
from __future__ import annotations

import numpy as np

from . import api, synth


def synth_sequence(seed: int, frames: int, pts_per_scan: int = 3000, raw_per_scan: int = 4000, extent: float = 400.0,
                   scan_radius: float = 60.0):
    """(list of static-point scans in the sensor frame [n_i,4], list of ground-truth poses T_w_s)."""
    rng = np.random.default_rng(seed)
    scene = synth.Scene(seed, extent=extent, n_walls=int(12 * (extent / 80.0) ** 2))
    poses = synth.trajectory(seed, frames)
    scans = []
    for T in poses:
        w = scene.sample(rng, raw_per_scan, centre=(T[0, 3], T[1, 3]), radius=scan_radius)
        s = synth.apply(np.linalg.inv(T), w)
        # the Doppler filter's job (static / dynamic split) is emulated by the generator's labels
        keep = rng.uniform(size=raw_per_scan) < pts_per_scan / raw_per_scan
        scans.append(np.ascontiguousarray(s[keep]))
    return scans, poses


def run_odometry(h: api.Icp4r, scans, opts: api.Opts, T_first=None, fused: bool = True):
    """Returns the list of estimated poses T_w_s (float64 4x4). fused: one icp4r_odometry_step call per frame
    (register + transform + Add_Points on the device); otherwise the three separate calls."""
    T = np.eye(4) if T_first is None else np.asarray(T_first, np.float64)
    poses = [T.copy()]
    if fused:
        h.map_build(np.zeros((0, 4), np.float32))
        h.odometry_step(scans[0], opts, T)
        for scan in scans[1:]:
            T, _res = h.odometry_step(scan, opts, T)
            poses.append(T.copy())
        return poses
    h.map_build(h.transform_points(T, scans[0]))
    for scan in scans[1:]:
        for i in range(16):
            opts.T0[i] = float(T.reshape(16)[i])
        T, res, _ = h.register_map(scan, opts)
        poses.append(T.copy())
        h.map_add_points(h.transform_points(T, scan), False)
    return poses


def run_odometry_raw(h: api.Icp4r, frames, opts: api.Opts, seed: int = 1, on_device=None):
    """BASELINE config 3 from RAW radar frames [n,5]: per frame the Doppler static-point filter
    (icp4r_doppler_static_points), then one icp4r_odometry_step (register against the growing map, transform, Add_Points).
    on_device: callable that moves one frame to the device (host-buffer flavour). Returns the estimated poses."""
    T = np.eye(4)
    poses = []
    h.map_build(np.zeros((0, 4), np.float32))
    for f, rec in enumerate(frames):
        if on_device is not None:
            rec = on_device(rec)
        static, _ = h.doppler_static_points(rec, 0, seed=seed + f)
        if f == 0:
            h.odometry_step(static, opts, T)
        else:
            T, _res = h.odometry_step(static, opts, T)
        poses.append(T.copy())
    return poses


# ---- scan-to-scan node ---------------------------------------------------------------------------------------------
def synth_radar_sequence(seed: int, frames: int, pts_per_frame: int = 1200, extent: float = 120.0, scan_radius: float = 50.0,
                         dynamic_frac: float = 0.1, fov_deg: float | None = None, max_range: float | None = None, forward: str = "x"):
    """(list of raw radar frames [n,5] x,y,z,intensity,doppler in the sensor frame, list of ground-truth poses).
    fov_deg / max_range: keep only the returns inside the sensor's field of view (|azimuth| <= fov_deg, range <=
    max_range in the sensor frame) like a forward-looking radar (the reference's sub-map is the same sector: 80 m,
    RADAR_RADIUS at radar_odometry.cpp:36); the scene is over-sampled so that about pts_per_frame returns remain.
    forward: the sensor axis the field of view is centred on. The scan-to-map node passes its yaw as the sector heading
    (radar_odometry.cpp:379,396) and KD_TREE::calc_heading measures headings from the +y axis (ikd_Tree.cpp:1434-1448:
    heading 0 = +y, -90 = +x), i.e. that node's sector looks along the body's +y axis: use forward="y" for its flow."""
    rng = np.random.default_rng(seed)
    scene = synth.Scene(seed, extent=extent, n_walls=int(12 * (extent / 80.0) ** 2))
    poses = synth.trajectory(seed, frames)
    out = []
    crop = fov_deg is not None or max_range is not None
    for f, T in enumerate(poses):
        n_draw = pts_per_frame * (6 if crop else 1)
        w = scene.sample(rng, n_draw, centre=(T[0, 3], T[1, 3]), radius=scan_radius)
        s = synth.apply(np.linalg.inv(T), w)
        if crop:
            keep = np.ones(len(s), bool)
            if max_range is not None:
                keep &= np.linalg.norm(s[:, :3], axis=1) <= max_range
            if fov_deg is not None:
                az = np.arctan2(s[:, 1], s[:, 0]) if forward == "x" else np.arctan2(-s[:, 0], s[:, 1])
                keep &= np.abs(np.degrees(az)) <= fov_deg
            s = np.ascontiguousarray(s[keep][:pts_per_frame])
        nxt = poses[min(f + 1, frames - 1)]
        prv = poses[max(f - 1, 0)]
        v_world = (nxt[:3, 3] - prv[:3, 3]) / (0.1 * max(1, min(f + 1, frames - 1) - max(f - 1, 0)))   # 10 Hz frames
        v_sensor = T[:3, :3].T @ v_world
        vr, _dyn = synth.doppler(rng, s, v_sensor, dynamic_frac)
        out.append(synth.radar_frame_bin(s, vr))
    return out, poses


def run_scan_to_scan(h: api.Icp4r, frames, opts: api.Opts, batched: bool = True, seed: int = 1):
    """The scan-to-scan node over the C ABI (what /root/reference/src/iterative_closest_point.cpp:300-560 does per
    frame): Doppler static-point filter on every frame (fitSineRansac + split, :387-403), ICP of the current static
    points (source) against the previous frame's (target) (:510-514), pose chained by RIGHT multiplication
    `currOdom = currOdom * icp_result` (:552). batched: all frame pairs of the recording in one
    icp4r_register_batch call (replay); otherwise one icp4r_register per frame (live).
    Returns (poses [frames][4,4], ego velocities [frames][3], per-frame results)."""
    statics, vel = [], []
    for f, rec in enumerate(frames):
        mask, dr = h.doppler_filter(rec, 0, seed=seed + f)
        m = np.asarray(mask.cpu() if hasattr(mask, "cpu") else mask).astype(bool)
        r = np.asarray(rec.cpu() if hasattr(rec, "cpu") else rec)
        statics.append(np.ascontiguousarray(r[m][:, :4], np.float32))
        vel.append(np.array(list(dr.velocity)))
    n = len(frames)
    if batched and n > 1:
        src = np.concatenate(statics[1:])
        tgt = np.concatenate(statics[:-1])
        so = np.concatenate([[0], np.cumsum([len(s) for s in statics[1:]])]).astype(np.int32)
        to = np.concatenate([[0], np.cumsum([len(s) for s in statics[:-1]])]).astype(np.int32)
        Ts, res = h.register_batch(src, so, tgt, to, opts)
    else:
        Ts, res = [], []
        for f in range(1, n):
            T, r, _ = h.register(statics[f], statics[f - 1], opts)
            Ts.append(T)
            res.append(r)
    cur = np.eye(4)
    poses = [cur.copy()]
    for T in Ts:
        cur = cur @ np.asarray(T, np.float64).reshape(4, 4)
        poses.append(cur.copy())
    return poses, vel, res


# ---- the scan-to-map node's own per-frame flow ---------------------------------------------------------------------
def yaw_deg(T) -> float:
    """heading as the reference computes it: R2rpy(R)(2) in degrees (/root/reference/src/radar_odometry.cpp:120-135,379)"""
    return float(np.degrees(np.arctan2(T[1, 0], T[0, 0])))


def run_reference_flow(h: api.Icp4r, frames, opts: api.Opts, radius: float = 80.0, leaf: float = 0.5, seed: int = 1, vg_out=None,
                       on_device=None, priors=None):
    """What /root/reference/src/radar_odometry.cpp:328,380-429 does per radar frame, with every step on the device:

      static points of the raw frame (ego-velocity / Doppler filter, :328)   -> icp4r_doppler_static_points
      pointAssociateToMap with the current pose (:384-389)                    -> icp4r_transform_points
      ikd_Tree.Add_Points(scan_map, false) (:390; Build on the first frame)   -> icp4r_map_add_points / icp4r_map_build
      ikd_Tree.Sector_Search(p_now, 80 m, heading) (:392-396)                 -> icp4r_map_sector (indices stay on the device)
      FastGICP align of the scan against the sub-map (:399-405)               -> icp4r_register_submap (ICP4R_GICP)
      currOdom = icp_result * currOdom (:412)
      VoxelGrid 0.5 m over the whole accumulated map (:426-429)               -> icp4r_voxel_grid(NULL)

    frames: raw radar records [n,5] (x, y, z, intensity, doppler), numpy or torch CUDA tensors; on_device: callable that
    moves one frame to the device (the bench's host-buffer flavour passes a pinned-memory copy here so that the raw frame
    is the only thing that crosses the bus). priors: the pose q_w_curr / t_w_curr the node places every scan with before
    aligning it — in the reference it comes from the odometry / ground-truth queue (radar_odometry.cpp:352-379); None: the
    chained estimate itself. Returns (currOdom per frame, points of the last down-sampled map)."""
    odom = np.eye(4)
    poses = []
    n_ds = 0
    for f, rec in enumerate(frames):
        if on_device is not None:
            rec = on_device(rec)
        T = odom if priors is None else np.asarray(priors[f], np.float64)
        static, _dres = h.doppler_static_points(rec, 0, seed=seed + f)
        scan_w = h.transform_points(T, static)
        if f == 0:
            h.map_build(scan_w)
            if priors is not None:
                odom = T.copy()
        else:
            h.map_add_points(scan_w, False)
            if isinstance(scan_w, np.ndarray):
                idx = h.map_sector(T[:3, 3], radius, yaw_deg(T))
            else:
                idx = h.map_sector_dev(T[:3, 3], radius, yaw_deg(T))
            D, _res = h.register_submap(scan_w, idx, opts)
            odom = D @ (odom if priors is None else T)
        poses.append(odom.copy())
        ds = h.voxel_grid(None, leaf, out=vg_out)
        n_ds = int(ds.shape[0])
    return poses, n_ds
