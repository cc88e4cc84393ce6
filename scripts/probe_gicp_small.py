"""GPU probe (for an ncu launch list): three GICP registrations, map-resident target, C2 shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
import bench
mp, scans = bench.make_c2()
h = pkg.Icp4r(0)
h.map_build(torch.from_numpy(mp).cuda())
d = torch.from_numpy(scans[0]).cuda()
o = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=64, max_corr_dist=0.0, early_exit=1)
for _ in range(3):
    T, res, _ = h.register_map(d, o)
print("iterations", res.iterations)
