"""2-GPU probe (torchrun): what does the cross-rank exchange cost per iteration? Every rank holds the FULL C5 map and
owns every point (slab = the whole axis), so the work per rank equals the single-GPU run and the difference to
register_map is the exchange + rank skew (the summed pose is meaningless here — timing only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from icp4r_loader import pkg
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))); torch.cuda.set_device(dev)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
mp, scans = bench.make_c5(M)
h = pkg.Icp4r(dev.index)
st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
h.map_build(torch.from_numpy(mp).to(dev))
d = torch.from_numpy(scans[0]).to(dev)
uid = [pkg.Icp4r.shard_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
h.shard_init(uid[0], rank, world)
def ev(fn, reps=20):
    with torch.cuda.stream(st):
        for _ in range(3): fn()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): fn()
        e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=20, max_corr_dist=2.0)
t_single = ev(lambda: h.register_map(d, o))
t_nccl = ev(lambda: h.register_sharded(d, o, 0, -1e30, 1e30))
hs = [None] * world
dist.all_gather_object(hs, h.shard_ipc_export())
h.shard_ipc_import(hs, rank, world)
t_peer = ev(lambda: h.register_sharded(d, o, 0, -1e30, 1e30))
if rank == 0:
    print(f"M={M}: register_map {t_single:.3f} ms | sharded, every point owned: NCCL {t_nccl:.3f} ms (+{(t_nccl - t_single) / 20 * 1e3:.1f} us/iter), "
          f"in-kernel peer exchange {t_peer:.3f} ms (+{(t_peer - t_single) / 20 * 1e3:.1f} us/iter)")
dist.barrier(); dist.destroy_process_group()
