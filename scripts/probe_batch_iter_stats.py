import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from icp4r_loader import pkg
import bench
mp, scans = bench.make_c2()
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
h.map_build(mp)
cat = torch.from_numpy(np.concatenate(scans)).to(dev)
off = (np.arange(17) * 4096).astype(np.int32)
h.set_stats(True)
prev = [0, 0, 0]
for it in range(0, 21):
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=it, max_corr_dist=2.0)
    h.get_stats()
    h.register_map_batch(cat, off, o)
    st = h.get_stats()[:3]
    d = [st[j] - prev[j] for j in range(3)]
    print(f"iterations {it:2d}: searches {st[0]:8d} (+{d[0]:7d})  dist evals {st[1]:10d} (+{d[1]:9d}, {d[1] / max(d[0], 1):6.1f} per search)  settled {st[2]:8d} (+{d[2]:7d})", flush=True)
    prev = st
