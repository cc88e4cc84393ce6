"""GPU probe (target for ncu): run one of the bench's workloads a few times so that a launch of its dominant kernel can
be captured. usage: probe_ncu_target.py c1|c2single|c2batch|c3|c4|c5 [size]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icp4r_loader import pkg
import bench

which = sys.argv[1]
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
if which == "c1":
    a, b, _ = pkg.synth.frame_pair(1001, 1024)
    da, db = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=30)
    for _ in range(4):
        T, r, _ = h.register(da, db, o)
elif which in ("c2single", "c2batch"):
    mp, scans = bench.make_c2()
    h.map_build(mp)
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
    if which == "c2single":
        d = torch.from_numpy(scans[0]).to(dev)
        for _ in range(4):
            T, r, _ = h.register_map(d, o)
    else:
        cat = torch.from_numpy(np.concatenate(scans)).to(dev)
        off = (np.arange(17) * 4096).astype(np.int32)
        for _ in range(3):
            T, r = h.register_map_batch(cat, off, o)
elif which == "c3":
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    raw, _ = pkg.pipeline.synth_radar_sequence(1003, frames, pts_per_frame=4000, extent=400.0, scan_radius=60.0)
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
    d = [torch.from_numpy(f).to(dev) for f in raw]
    pkg.pipeline.run_odometry_raw(h, d, o)
    last, _ = h.doppler_static_points(d[-1], 0, seed=frames)
    for _ in range(4):
        T, r, _ = h.register_map(last, o)
elif which == "c4":
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    src, tgt, off = bench.make_c4(pairs)
    o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=30)
    ds, dt, do = torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev), torch.from_numpy(off).to(dev)
    for _ in range(3):
        T, r = h.register_batch(ds, do, dt, do, o)
elif which == "c5":
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
    mp, scans = bench.make_c5(m)
    h.map_build(torch.from_numpy(mp).to(dev))
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
    d = torch.from_numpy(scans[0]).to(dev)
    for _ in range(4):
        T, r, _ = h.register_map(d, o)
h.synchronize()
print("ok", which)
