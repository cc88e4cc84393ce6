"""GPU probe: kNN time for queries far from the map's points (ring cap -> coarse-table bound -> one wide pass)."""
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
def ev(q, k, gate, reps=10):
    d = torch.from_numpy(np.ascontiguousarray(q)).to(dev)
    out = (torch.empty((len(q), k), dtype=torch.int32, device=dev), torch.empty((len(q), k), dtype=torch.float32, device=dev),
           torch.empty(len(q), dtype=torch.int32, device=dev))
    with torch.cuda.stream(st):
        for _ in range(2): h.map_knn(d, k, gate, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): h.map_knn(d, k, gate, out=out)
        e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
rng = np.random.default_rng(0)
base = scans[0].copy()
for name, q in (("scan itself", base),
                ("scan lifted 6 m above the map", base + np.array([0, 0, 6.0, 0], np.float32)),
                ("scan shifted 30 m sideways", base + np.array([30.0, 0, 0, 0], np.float32)),
                ("scan 300 m away", base + np.array([300.0, 0, 0, 0], np.float32))):
    print(f"{name:32s} k=5 ungated {ev(q, 5, 0.0):9.1f} us   k=1 ungated {ev(q, 1, 0.0):9.1f} us   k=5 gate 20 m {ev(q, 5, 20.0):9.1f} us")
