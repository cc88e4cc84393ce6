"""GPU probe: C5 per-iteration time vs the map's cell size (0 = the library's own occupancy-based choice)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
import bench
M = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
mp, scans = bench.make_c5(M)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream()
d_map = torch.from_numpy(mp).to(dev)
d = torch.from_numpy(scans[0]).to(dev)
h = pkg.Icp4r(0); h.set_stream(st.cuda_stream)
for cell in (0.0, 0.4, 0.5, 0.6, 0.7, 0.85, 1.0, 1.3):
    h.map_build(d_map, cell)
    res = []
    for iters in (4, 20):
        o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=iters, max_corr_dist=2.0)
        with torch.cuda.stream(st):
            for _ in range(3): h.register_map(d, o)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(10): h.register_map(d, o)
            e1.record(st)
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 10)
    print(f"cell {cell:4.2f}: 4 it {res[0]:.3f} ms, 20 it {res[1]:.3f} ms -> {(res[1] - res[0]) / 16 * 1e3:.1f} us / iteration", flush=True)
