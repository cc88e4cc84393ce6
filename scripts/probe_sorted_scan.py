"""GPU probe: does the order of the scan's points matter? register the C2 batch with scans as generated and with scans
sorted by the map cell their initial position falls in (x fastest, like the map's own order)."""
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
o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
def sort_scan(s, cell):
    ijk = np.floor(s[:, :3] / cell).astype(np.int64)
    key = (ijk[:, 2] * 100000 + ijk[:, 1]) * 100000 + ijk[:, 0]
    return np.ascontiguousarray(s[np.argsort(key, kind="stable")])
B = 16
for name, cell in (("as generated", None), ("sorted, 0.5 m cells", 0.5), ("sorted, 2 m cells", 2.0), ("sorted, 8 m cells", 8.0)):
    ss = [scans[i % 8] if cell is None else sort_scan(scans[i % 8], cell) for i in range(B)]
    S = torch.from_numpy(np.concatenate(ss)).to(dev)
    off = (np.arange(B + 1) * 4096).astype(np.int32)
    with torch.cuda.stream(st):
        for _ in range(3): h.register_map_batch(S, off, o)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(20): h.register_map_batch(S, off, o)
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    d1 = torch.from_numpy(ss[0]).to(dev)
    with torch.cuda.stream(st):
        for _ in range(3): h.register_map(d1, o)
        e0.record(st)
        for _ in range(20): h.register_map(d1, o)
        e1.record(st)
    torch.cuda.synchronize()
    print(f"{name:22s}: batch of 16 {ms:.3f} ms = {B / ms * 1e3:7.0f} registrations/s;  single scan {e0.elapsed_time(e1) / 20:.4f} ms")
