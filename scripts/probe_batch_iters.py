import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from icp4r_loader import pkg
import bench
mp, scans = bench.make_c2()
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
h.map_build(mp)
cat = torch.from_numpy(np.concatenate(scans)).to(dev)
off = (np.arange(17) * 4096).astype(np.int32)
prev = 0.0
for it in (0, 1, 2, 3, 4, 6, 8, 10, 14, 20):
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=it, max_corr_dist=2.0)
    ts = []
    for r in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); h.register_map_batch(cat, off, o); e1.record(st); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts[2:]))
    print(f"iterations {it:2d}: {t:.3f} ms  (+{t - prev:.3f})", flush=True)
    prev = t
