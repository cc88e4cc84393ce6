"""GPU probe: icp4r_voxel_grid time vs cloud size (leaf 0.5 m, uniform 400 x 400 x 20 m volume), device-resident in/out."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
for n in (100_000, 1_000_000, 6_000_000, 20_000_000):
    g = torch.Generator(device="cuda").manual_seed(n)
    p = torch.rand((n, 4), generator=g, device=dev)
    p[:, 0] = (p[:, 0] - 0.5) * 400; p[:, 1] = (p[:, 1] - 0.5) * 400; p[:, 2] = (p[:, 2] - 0.5) * 20
    with torch.cuda.stream(st):
        for _ in range(2): out = h.voxel_grid(p, 0.5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5): out = h.voxel_grid(p, 0.5)
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"n {n:9d} -> {out.shape[0]:8d} leaves: {ms:7.3f} ms  ({n * 16 / ms / 1e6:7.1f} GB/s of input)")
