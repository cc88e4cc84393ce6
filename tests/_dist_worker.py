"""Worker for the multi-process tests (launched by tests/test_distributed_cpu.py and tests/test_gpu_sharded.py).

mode cpu : gloo, no GPU. Checks the sharding math with the ORACLE standing in for the device kernel:
           sum over ranks of (accumulators of the source points a rank owns, searched in its slab+halo)
           == accumulators over the whole map, and the per-rank pair ranges tile the batch.
mode gpu : one rank per GPU, icp4r_register_sharded first with the NCCL all-reduce, then with the in-kernel peer-memory
           exchange (icp4r_shard_ipc_*); the sharded pose must equal the single-GPU pose on the whole map and the oracle's.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icp4r_loader import pkg  # noqa: E402
import oracle as O  # noqa: E402


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    gate = 2.0
    scan, mp, _ = pkg.synth.scan_to_map(77, 1500, 60000, extent=60.0, scan_radius=50.0)
    bounds = pkg.shard.slab_bounds(mp[:, 0], world)
    mine, lo, hi, gidx = pkg.shard.slab_of_rank(mp, rank, world, axis=0, halo=gate, bounds=bounds)
    oo = O.default_opts(residual=O.P2PLANE_KNN, k=5, max_iterations=6, max_corr_dist=gate)

    if mode == "cpu":
        T = pkg.synth.se3(0.01, 0, 0, (0.2, -0.1, 0.0))
        q32, _ = O.transform(T, scan)
        own = pkg.shard.owner_of(q32[:, 0], bounds)
        cnt = torch.tensor([int((own == rank).sum())])
        dist.all_reduce(cnt)
        assert int(cnt.item()) == len(scan), "every source point must have exactly one owner"
        sel = scan[own == rank]
        acc_local = np.zeros(O.ACC_LEN)
        if len(sel):
            acc_local, idx_local, used = O.accumulate(sel, mine, oo, T)
            # the halo guarantee: searching the slab finds exactly the neighbours the whole map would
            _, idx_full, _ = O.accumulate(sel, mp, oo, T)
            assert (np.where(idx_local >= 0, gidx[np.maximum(idx_local, 0)], -1) == idx_full).all()
        t = torch.from_numpy(acc_local.copy())
        dist.all_reduce(t)  # the 29-double all-reduce of the device path, on gloo
        acc_full, _, _ = O.accumulate(scan, mp, oo, T)
        assert np.allclose(t.numpy()[:29], acc_full[:29], rtol=1e-9, atol=1e-9), np.abs(t.numpy()[:29] - acc_full[:29]).max()
        # batched pairs: the ranks' ranges tile [0, P) without overlap
        r = torch.tensor(list(pkg.shard.pair_range(65536 + 7, rank, world)))
        allr = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
        dist.all_gather(allr, r)
        flat = [int(x) for t2 in allr for x in t2]
        assert flat[0] == 0 and flat[-1] == 65536 + 7 and all(flat[2 * i + 1] == flat[2 * i + 2] for i in range(world - 1))
        if rank == 0:
            print("DIST-CPU-OK")
    else:
        dev = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(dev)
        h = pkg.Icp4r(dev)
        uid = [pkg.Icp4r.shard_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.shard_init(uid[0], rank, world)
        h.map_build(mine)
        o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=6, max_corr_dist=gate)
        T, res = h.register_sharded(scan, o, 0, lo, hi)
        Ts = [None] * world
        dist.all_gather_object(Ts, T)
        for other in Ts:
            assert np.array_equal(other, T), "every rank must return the same pose"
        if rank == 0:
            h2 = pkg.Icp4r(dev)
            h2.map_build(mp)
            T1, r1, _ = h2.register_map(scan, o)
            To, ro, _ = O.register(scan, mp, oo)
            assert np.abs(T - T1).max() < 1e-9, np.abs(T - T1).max()
            assert np.abs(T - To).max() < 1e-8, np.abs(T - To).max()
            assert res.n_corr == r1.n_corr == ro.n_corr and res.n_fitness == r1.n_fitness
            assert abs(res.fitness - r1.fitness) < 1e-9 * max(r1.fitness, 1)
            print("DIST-GPU-OK", res.n_corr)
        # fused flavour: the cross-rank sum runs inside the iteration kernel over peer memory (CUDA IPC), no NCCL call
        mine_h = h.shard_ipc_export()
        hs = [None] * world
        dist.all_gather_object(hs, mine_h)
        h.shard_ipc_import(hs, rank, world)
        for rep, oe in enumerate([o, o, pkg.default_opts(residual=pkg.P2P_SVD, k=1, max_iterations=40, max_corr_dist=gate, early_exit=1), o]):
            dist.barrier()
            Tf, rf = h.register_sharded(scan, oe, 0, lo, hi)
            Ts = [None] * world
            dist.all_gather_object(Ts, (Tf, rf.iterations, rf.n_corr))
            for other in Ts:
                assert np.array_equal(other[0], Tf) and other[1:] == (rf.iterations, rf.n_corr), "ranks disagree (fused exchange)"
            if oe is o:
                assert np.abs(Tf - T).max() < 1e-10, (rep, np.abs(Tf - T).max())
                assert rf.n_corr == res.n_corr and rf.n_fitness == res.n_fitness
            elif rank == 0:
                T1e, r1e, _ = h2.register_map(scan, oe)
                assert np.abs(Tf - T1e).max() < 1e-9 and rf.iterations == r1e.iterations < 40, (rf.iterations, r1e.iterations)
        if rank == 0:
            print("DIST-GPU-FUSED-OK")
        h.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
