"""GPU probe: which queries make the kNN kernel slow? time subsets (dev aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
import bench
mp, scans = bench.make_c2()
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
h.map_build(mp)
scan = scans[0]
idx, d2, found = h.map_knn(scan, 5, 2.0)
print("found histogram", np.bincount(found, minlength=6), "kth dist mean", np.sqrt(d2[found == 5, 4]).mean())
def ev_time(q, k, gate, reps=30):
    d = torch.from_numpy(np.ascontiguousarray(q)).to(dev)
    out = (torch.empty((len(q), k), dtype=torch.int32, device=dev), torch.empty((len(q), k), dtype=torch.float32, device=dev),
           torch.empty(len(q), dtype=torch.int32, device=dev))
    with torch.cuda.stream(st):
        for _ in range(3): h.map_knn(d, k, gate, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): h.map_knn(d, k, gate, out=out)
        e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
full = scan[found == 5]
sparse = scan[found < 5]
print("n full", len(full), "n sparse", len(sparse))
for name, q in (("all", scan), ("full5 x400", full[:400]), ("sparse x400", sparse[:400]), ("full5 all", full), ("1 query", scan[:1]), ("32 queries", scan[:32])):
    print(f"{name:14s} n={len(q):5d}  k=5 gate2: {ev_time(q,5,2.0):6.1f} us   k=1 ungated: {ev_time(q,1,0.0):6.1f} us")
