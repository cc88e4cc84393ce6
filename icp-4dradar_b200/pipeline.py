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
