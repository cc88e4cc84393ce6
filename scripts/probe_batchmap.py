"""GPU probe: throughput of icp4r_register_map_batch vs the number of scans per call (C2 shapes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
import bench
mp, scans = bench.make_c2()
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
h.map_build(mp)
o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
for B in (1, 2, 4, 8, 16, 32, 64):
    S = torch.from_numpy(np.concatenate([scans[i % 8] for i in range(B)])).to(dev)
    off = (np.arange(B + 1) * 4096).astype(np.int32)
    with torch.cuda.stream(st):
        for _ in range(3): h.register_map_batch(S, off, o)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        reps = 10
        for _ in range(reps): h.register_map_batch(S, off, o)
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"B={B:3d}: {ms:8.3f} ms per call  {B / ms * 1e3:9.1f} registrations/s")
