"""GPU probe: split the iteration time into kNN / accumulate / solve by timing variants (dev aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
import bench

mp, scans = bench.make_c2()
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
st = torch.cuda.Stream()
h.set_stream(st.cuda_stream)
h.map_build(mp)
d = [torch.from_numpy(s).to(dev) for s in scans]

def ev_time(fn, reps=20):
    with torch.cuda.stream(st):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): fn()
        e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for k, gate in ((1, 0.0), (1, 2.0), (5, 2.0), (5, 0.0)):
    out = (torch.empty((4096, k), dtype=torch.int32, device=dev), torch.empty((4096, k), dtype=torch.float32, device=dev),
           torch.empty(4096, dtype=torch.int32, device=dev))
    print(f"map_knn k={k} gate={gate}: {ev_time(lambda: h.map_knn(d[0], k, gate, out=out)):.1f} us")
h.set_profiling(True)
for name, kind, k, gate in (("P2PLANE", pkg.P2PLANE_KNN, 5, 2.0), ("P2P_SVD", pkg.P2P_SVD, 1, 2.0), ("P2P_GN", pkg.P2P_GN, 1, 2.0), ("P2LINE", pkg.P2LINE, 2, 2.0)):
    o = pkg.default_opts(residual=kind, k=k, max_iterations=20, max_corr_dist=gate)
    with torch.cuda.stream(st):
        for _ in range(3):
            h.register_map(d[0], o)
        pr = h.last_profile()
    print(name, "iter us:", np.round(pr[:20].mean() * 1e3, 1), "min", np.round(pr[:20].min() * 1e3, 1), "fitness us:", np.round(pr[20] * 1e3, 1))
h.set_profiling(False)
for name, kind, k, gate in (("P2PLANE", pkg.P2PLANE_KNN, 5, 2.0), ("P2P_SVD", pkg.P2P_SVD, 1, 2.0)):
    for it in (1, 20, 40):
        o = pkg.default_opts(residual=kind, k=k, max_iterations=it, max_corr_dist=gate)
        print(name, "graph iters", it, f"{ev_time(lambda: h.register_map(d[0], o)):.1f} us per registration")
