"""GPU probe: GICP registration time on the C2 shapes (4096-pt scan vs 200k-pt map), and against a sector sub-map."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
import bench
mp, scans = bench.make_c2()
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
h.map_build(torch.from_numpy(mp).to(dev))
d = [torch.from_numpy(s).to(dev) for s in scans]
def ev(fn, reps=10):
    with torch.cuda.stream(st):
        for _ in range(3): r = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): r = fn()
        e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r
for ee, iters in ((0, 20), (1, 64)):
    o = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=iters, max_corr_dist=2.0, early_exit=ee)
    ms, (T, res, _) = ev(lambda: h.register_map(d[0], o))
    print(f"GICP map-resident early_exit={ee} max_it={iters}: {ms:.3f} ms  iterations {res.iterations} converged {res.converged} n_corr {res.n_corr}")
    h.set_profiling(1)
    h.register_map(d[0], o)
    print("   profile ms:", np.round(h.last_profile()[:12], 3))
    h.set_profiling(0)
o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
ms, _ = ev(lambda: h.register_map(d[0], o))
print(f"P2PLANE_KNN 20 it: {ms:.3f} ms")
# the reference's own flow: pair registration (scan vs extracted sub-map), target index + normals built every call
sub = mp[np.linalg.norm(mp[:, :2], axis=1) < 80.0]
dsub = torch.from_numpy(np.ascontiguousarray(sub)).to(dev)
o = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=64, max_corr_dist=0.0, early_exit=1)
ms, (T, res) = ev(lambda: h.register(d[0], dsub, o)[:2], reps=5)
print(f"GICP pair (scan vs {len(sub)}-pt sub-map, build + normals every call): {ms:.3f} ms  iterations {res.iterations}")
import time
def wall(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r
h2 = pkg.Icp4r(0)
for rep in range(3):
    t_build, _ = wall(lambda: h2.map_build(dsub))
    o = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=64, max_corr_dist=0.0, early_exit=1)
    t_first, (T, res, _) = wall(lambda: h2.register_map(d[0], o))
    t_second, _ = wall(lambda: h2.register_map(d[0], o))
    o1 = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=1, max_corr_dist=0.0, early_exit=1)
    t_one, _ = wall(lambda: h2.register_map(d[0], o1))
    print(f"sub-map build {t_build:.3f} ms; GICP first call (target normals) {t_first:.3f} ms; second call {t_second:.3f} ms ({res.iterations} its); 1-iteration call {t_one:.3f} ms")
