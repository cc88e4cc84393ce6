"""GPU probe: cost of Build / Add_Points(false) as the map grows (C3 shape) and of a registration against it."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
s = pkg.synth
h = pkg.Icp4r(0)
rng = np.random.default_rng(1003)
sc = s.Scene(1003, extent=400.0, n_walls=200)
dev = torch.device("cuda", 0)
def t(fn, reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
for M in (100_000, 1_000_000, 6_000_000):
    mp = sc.sample(rng, M)
    d = torch.from_numpy(mp).to(dev)
    add = torch.from_numpy(sc.sample(rng, 3000, centre=(10.0, 5.0), radius=80.0)).to(dev)
    print(f"M={M}: map_build {t(lambda: h.map_build(d)):.2f} ms", end="  ")
    h.map_build(d)
    t0 = time.perf_counter(); h.map_add_points(add, False); torch.cuda.synchronize()
    print(f"add_points(3000) {(time.perf_counter()-t0)*1e3:.2f} ms", end="  ")
    scan = s.apply(np.linalg.inv(s.random_small_se3(rng)), sc.sample(rng, 3000, centre=(10.0, 5.0), radius=80.0))
    ds = torch.from_numpy(scan).to(dev)
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
    print(f"register_map {t(lambda: h.register_map(ds, o), 5):.3f} ms")
