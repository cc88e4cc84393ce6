"""GPU probe: where a frame of the node's own flow (pipeline.run_reference_flow) spends its time: per-call wall times
(stream-synchronised) at a few frames of the sequence. usage: probe_c3ref.py [frames]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icp4r_loader import pkg

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 120
h = pkg.Icp4r(0)
frames, gt = pkg.pipeline.synth_radar_sequence(1003, nf, pts_per_frame=4000, extent=400.0, scan_radius=90.0, fov_deg=55.0, max_range=78.0, forward="y")
o = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=64, early_exit=1, max_corr_dist=0.0)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(f).to(dev) for f in frames]
vg_out = torch.empty((sum(len(f) for f in frames), 4), dtype=torch.float32, device=dev)


def T(fn):
    h.synchronize()
    t0 = time.perf_counter()
    r = fn()
    h.synchronize()
    return r, 1e3 * (time.perf_counter() - t0)


for f, rec in enumerate(d):
    P = gt[f]
    (static, _), t_dop = T(lambda: h.doppler_static_points(rec, 0, seed=1 + f))
    scan_w, t_tr = T(lambda: h.transform_points(P, static))
    if f == 0:
        h.map_build(scan_w)
        continue
    _, t_add = T(lambda: h.map_add_points(scan_w, False))
    idx, t_sec = T(lambda: h.map_sector_dev(P[:3, 3], 80.0, pkg.pipeline.yaw_deg(P)))
    (D, res), t_reg = T(lambda: h.register_submap(scan_w, idx, o))
    ds, t_vg = T(lambda: h.voxel_grid(None, 0.5, out=vg_out))
    if f % 20 == 0 or f == nf - 1:
        print(f"frame {f}: static {len(static)} map {h.map_size()[0]} sub-map {len(idx)} | doppler {t_dop:.2f} transform {t_tr:.2f} add {t_add:.2f} sector {t_sec:.2f} "
              f"register_submap {t_reg:.2f} (iterations {res.iterations}, converged {res.converged}) voxel_grid {t_vg:.2f} ms | |D - I| {np.abs(D - np.eye(4)).max():.2e}", flush=True)
