"""GPU probe (target for ncu): three icp4r_register_map_batch calls of 16 C2 scans. With the handle's profiling mode
the per-iteration kernel times of a single-scan registration are printed as well (argument: prof)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icp4r_loader import pkg
import bench

mp, scans = bench.make_c2()
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
h.map_build(mp)
o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
cat = torch.from_numpy(np.concatenate(scans)).to(dev)
off = (np.arange(17) * 4096).astype(np.int32)
for _ in range(3):
    T, r = h.register_map_batch(cat, off, o)
h.synchronize()
print("ok", T[0, 0, 3] if hasattr(T, "shape") else T)
if len(sys.argv) > 1 and sys.argv[1] == "prof":
    d = torch.from_numpy(scans[0]).to(dev)
    h.set_profiling(True)
    for _ in range(3):
        h.register_map(d, o)
    print("per-launch ms (single scan):", " ".join(f"{1e3 * x:.1f}" for x in h.last_profile()), "us")
