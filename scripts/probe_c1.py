"""GPU probe: latency of one pair registration (icp4r_register: transient target index + loop) at C1 / scan-to-scan sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
for n in (1024, 2048, 4096):
    src, tgt, _ = pkg.synth.frame_pair(1001, n)
    ds, dt = torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev)
    for name, o in (("P2P_SVD 30 it ungated (C1)", pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=30)),
                    ("P2P_SVD 10 it ungated (PCL default)", pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=10)),
                    ("P2PLANE k=5 20 it gate 2 m", pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0))):
        with torch.cuda.stream(st):
            for _ in range(3): h.register(ds, dt, o)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(20): h.register(ds, dt, o)
            e1.record(st)
        torch.cuda.synchronize()
        print(f"n = m = {n}: {name:38s} {e0.elapsed_time(e1) / 20:.3f} ms")
