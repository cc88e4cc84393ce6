"""1-GPU probe: cost of the sharded code path without any peer (one-rank communicator): ownership pre-pass + the
exchange's fences and flag round trip through local memory."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
import bench
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
mp, scans = bench.make_c5(M)
dev = torch.device("cuda", 0)
h = pkg.Icp4r(0)
st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
h.map_build(torch.from_numpy(mp).to(dev))
d = torch.from_numpy(scans[0]).to(dev)
def ev(fn, reps=20):
    with torch.cuda.stream(st):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): fn()
        e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
t0 = ev(lambda: h.register_map(d, o))
t1 = ev(lambda: h.register_sharded(d, o, 0, -1e30, 1e30))   # no communicator: separate solve kernel per iteration
h.shard_ipc_import([h.shard_ipc_export()], 0, 1)
t2 = ev(lambda: h.register_sharded(d, o, 0, -1e30, 1e30))   # one-rank exchange inside the kernel
print(f"register_map {t0:.3f} ms | sharded path, no communicator {t1:.3f} ms (+{(t1 - t0) / 20 * 1e3:.1f} us/iter) | one-rank in-kernel exchange {t2:.3f} ms (+{(t2 - t0) / 20 * 1e3:.1f} us/iter)")
