"""GPU probe: run the stand-alone grid kNN kernel a few times (target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
import bench
mp, scans = bench.make_c2()
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
h.map_build(mp)
d = torch.from_numpy(scans[0]).to(dev)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
out = (torch.empty((4096, k), dtype=torch.int32, device=dev), torch.empty((4096, k), dtype=torch.float32, device=dev),
       torch.empty(4096, dtype=torch.int32, device=dev))
for _ in range(5):
    h.map_knn(d, k, 2.0, out=out)
torch.cuda.synchronize()
print("found mean", out[2].float().mean().item())
