"""GPU probe: scan-to-map registration times (C1, C2 single, C2 batch of 16, C5 at a reduced map) with the
keep-the-neighbours proof on and off (ICP4R_NO_LB), plus the work counters. usage: probe_map_kernels.py [c5_points]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from icp4r_loader import pkg
import bench

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
m5 = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
mp2, scans2 = bench.make_c2()
mp5, scans5 = bench.make_c5(m5)
c1 = [pkg.synth.frame_pair(1001 + i, 1024) for i in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, steps=30, warm=5):
    ts = []
    with torch.cuda.stream(stream):
        for i in range(warm + steps):
            flush.fill_(i & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn(i)
            e1.record(stream)
            ts.append((e0, e1))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ts[warm:]]))


for nolb in ("0", "1"):
    os.environ["ICP4R_NO_LB"] = nolb
    h = pkg.Icp4r(0)
    h.set_stream(stream.cuda_stream)
    o1 = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=30)
    d1 = [(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b, _ in c1]
    t_c1 = timed(lambda i: h.register(d1[i % 4][0], d1[i % 4][1], o1))
    o2 = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
    h.map_build(mp2)
    ds = [torch.from_numpy(s).to(dev) for s in scans2]
    t_c2 = timed(lambda i: h.register_map(ds[i % 16], o2))
    h.set_stats(True)
    T_single, r_single, _ = h.register_map(ds[0], o2)
    st = h.get_stats()
    h.set_stats(False)
    cat = torch.from_numpy(np.concatenate(scans2)).to(dev)
    off = (np.arange(17) * 4096).astype(np.int32)
    t_c2b = timed(lambda i: h.register_map_batch(cat, off, o2), steps=20)
    h.map_build(torch.from_numpy(mp5).to(dev))
    d5 = [torch.from_numpy(s).to(dev) for s in scans5]
    t_c5 = timed(lambda i: h.register_map(d5[i % 4], o2))
    T5, r5, _ = h.register_map(d5[0], o2)
    print(f"NO_LB={nolb}: C1 {t_c1:.3f} ms | C2 single {t_c2:.3f} ms | C2 batch16 {t_c2b:.3f} ms ({16e3 / t_c2b:.0f} reg/s) | C5({m5}) {t_c5:.3f} ms"
          f" | C2 stats searches {st[0]} cands {st[1]} settled {st[2]} | T2[0,3]={T_single[0, 3]:.12f} n_corr {r_single.n_corr} T5[0,3]={T5[0, 3]:.12f} n_corr {r5.n_corr}", flush=True)
    h.close()
