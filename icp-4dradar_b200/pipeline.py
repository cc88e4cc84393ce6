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


# ---- scan-to-scan node ---------------------------------------------------------------------------------------------
def synth_radar_sequence(seed: int, frames: int, pts_per_frame: int = 1200, extent: float = 120.0, scan_radius: float = 50.0,
                         dynamic_frac: float = 0.1):
    """(list of raw radar frames [n,5] x,y,z,intensity,doppler in the sensor frame, list of ground-truth poses)."""
    rng = np.random.default_rng(seed)
    scene = synth.Scene(seed, extent=extent, n_walls=int(12 * (extent / 80.0) ** 2))
    poses = synth.trajectory(seed, frames)
    out = []
    for f, T in enumerate(poses):
        w = scene.sample(rng, pts_per_frame, centre=(T[0, 3], T[1, 3]), radius=scan_radius)
        s = synth.apply(np.linalg.inv(T), w)
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
