"""GPU probe (build with `make -C icp-4dradar_b200/csrc XFLAGS=-DICP4R_PHASE_TIMING`): per-iteration phase cycles of
the fused iteration kernel on the C5 map for scans of 4096 / 8192 / 16384 points."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
import bench
M = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
mp, scans = bench.make_c5(M)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream()
h = pkg.Icp4r(0); h.set_stream(st.cuda_stream)
h.map_build(torch.from_numpy(mp).to(dev))
o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=4, max_corr_dist=2.0)
for n in (4096, 8192, 16384):
    d = torch.from_numpy(np.ascontiguousarray(scans[0][:n])).to(dev)
    h.set_profiling(0)
    for _ in range(3): h.register_map(d, o)
    h.set_profiling(1)
    sys.stderr.write(f"--- n = {n}\n"); sys.stderr.flush()
    h.register_map(d, o)
    print(n, "per-launch ms:", np.round(h.last_profile(), 4))
