"""GPU probe (target for ncu): a few launches of the batched-pairs kernel on C4-shaped pairs. usage: probe_c4_one.py [pairs] [launches]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from icp4r_loader import pkg
import bench

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
src, tgt, off = bench.make_c4(pairs)
dev = torch.device("cuda", 0)
d_src, d_tgt, d_off = torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev), torch.from_numpy(off).to(dev)
o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=30)
h = pkg.Icp4r(0)
for _ in range(reps):
    T, R = h.register_batch(d_src, d_off, d_tgt, d_off, o)
h.synchronize()
print("ok", float(T.sum().item()))
