"""GPU probe (one GPU): where a C5 registration's time goes — full map vs one rank's share of a 2-slab split
(half the map, half the source points owned), warm L2 vs flushed, by iteration count."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
import bench
M = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
mp, scans = bench.make_c5(M)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
d_scan = torch.from_numpy(scans[0]).to(dev)


def timeit(fn, reps=20, cold=False):
    tot = 0.0
    with torch.cuda.stream(st):
        for _ in range(3): fn()
        for _ in range(reps):
            if cold: flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st)
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
    return tot / reps


h = pkg.Icp4r(0); h.set_stream(st.cuda_stream)
h.map_build(torch.from_numpy(mp).to(dev))
bounds = pkg.shard.slab_bounds(mp[:, 0], 2)
mine, lo, hi, _ = pkg.shard.slab_of_rank(mp, 0, 2, axis=0, halo=2.0, bounds=bounds)
h2 = pkg.Icp4r(0); h2.set_stream(st.cuda_stream)
h2.map_build(torch.from_numpy(mine).to(dev))
for iters in (1, 5, 10, 20):
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=iters, max_corr_dist=2.0)
    row = []
    for cold in (False, True):
        row.append(timeit(lambda: h.register_map(d_scan, o), cold=cold))
        row.append(timeit(lambda: h2.register_sharded(d_scan, o, 0, lo, hi), cold=cold))
    print(f"iters {iters:2d}: full warm {row[0]:.3f}  half warm {row[1]:.3f}  full cold {row[2]:.3f}  half cold {row[3]:.3f} ms")
