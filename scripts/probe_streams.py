"""GPU probe (target for an ncu launch list): the streaming kernels once each — full grid build of a 20 M dense map, icp4r_voxel_grid
on 20 M points, incremental Add_Points of 3,000 points into a 3 M map. usage: probe_streams.py [points]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icp4r_loader import pkg

m = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
d = torch.from_numpy(pkg.synth.dense_map(1005, m)).to(dev)


FLUSH = os.environ.get("PROBE_FLUSH") == "1"   # write a 256 MiB buffer (> L2) before every timed call, like bench.py does
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if FLUSH else None


def ev(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if FLUSH:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        h.synchronize()
        e0.record()
        fn()
        h.synchronize()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


print(f"map_build({m}): {ev(lambda: h.map_build(d)):.3f} ms", flush=True)
g = torch.Generator(device="cuda").manual_seed(7)
p = torch.rand((m, 4), generator=g, device=dev)
p[:, 0] = (p[:, 0] - 0.5) * 400
p[:, 1] = (p[:, 1] - 0.5) * 400
p[:, 2] = (p[:, 2] - 0.5) * 20
print(f"voxel_grid({m}, 0.5): {ev(lambda: h.voxel_grid(p, 0.5)):.3f} ms", flush=True)
del p, d
s = pkg.synth
rng = np.random.default_rng(1003)
sc = s.Scene(1003, extent=400.0, n_walls=200)
h2 = pkg.Icp4r(0)
h2.map_build(torch.from_numpy(sc.sample(rng, 3_000_000)).to(dev))
ts = []
for f in range(8):
    w = torch.from_numpy(sc.sample(rng, 3000, centre=(10.0 + f, 5.0), radius=80.0)).to(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2.synchronize()
    e0.record()
    h2.map_add_points(w, False)
    h2.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("add_points(3000 into 3 M): " + " ".join(f"{t:.3f}" for t in ts) + " ms", flush=True)
