"""GPU probe: throughput of the stand-alone kNN kernel (no previous-iteration bound) for 4k .. 256k queries on the C2 map,
to compare with the fused iteration kernel's time for the same number of queries."""
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
for B in (1, 4, 16, 64):
    q = torch.from_numpy(np.concatenate([scans[i % 8] for i in range(B)])).to(dev)
    n = q.shape[0]
    for k, gate in ((5, 2.0), (1, 2.0)):
        out = (torch.empty((n, k), dtype=torch.int32, device=dev), torch.empty((n, k), dtype=torch.float32, device=dev),
               torch.empty(n, dtype=torch.int32, device=dev))
        with torch.cuda.stream(st):
            for _ in range(3): h.map_knn(q, k, gate, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(20): h.map_knn(q, k, gate, out=out)
            e1.record(st)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"queries {n:7d} k={k}: {us:8.1f} us  = {n / us:7.1f} M queries/s")
