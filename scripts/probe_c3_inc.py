"""GPU probe: per-frame cost split of the odometry loop at a given map size: register, transform, incremental add."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
s = pkg.synth
h = pkg.Icp4r(0)
rng = np.random.default_rng(1003)
sc = s.Scene(1003, extent=400.0, n_walls=200)
dev = torch.device("cuda", 0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
mp = sc.sample(rng, M)
h.map_build(torch.from_numpy(mp).to(dev))
o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
def wall(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r
for f in range(8):
    w = sc.sample(rng, 3000, centre=(10.0 + f, 5.0), radius=80.0)
    scan = s.apply(np.linalg.inv(s.random_small_se3(rng)), w)
    ds = torch.from_numpy(scan).to(dev)
    t_reg, (T, res, _) = wall(lambda: h.register_map(ds, o))
    t_tr, wpts = wall(lambda: h.transform_points(T, ds))
    t_add, _ = wall(lambda: h.map_add_points(wpts, False))
    print(f"frame {f}: register {t_reg:.3f} ms  transform {t_tr:.3f} ms  add_points {t_add:.3f} ms  (map {h.map_size()[0]})", flush=True)
