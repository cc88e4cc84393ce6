#!/usr/bin/env python
"""bench.py — registrations/s and NN queries/s of the registration hot path on B200, with the reference's
CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4]

Workload (default, BASELINE.json configs[1] = "C2"): one step = ONE scan-to-map registration, a 4,096-point
synthetic 4D-radar scan against a resident 200,000-point map, point-to-plane with k = 5 neighbours
(LidarPlaneNormFactor), 20 iterations, 2.0 m gate.  The map is built once and stays resident in HBM like the
reference's ikd-Tree stays resident in RAM; scans rotate through a pool of 8 different scans.
  value : registrations/s with the scan already in HBM (device pointer through the C ABI)
  e2e   : the same through the C ABI with a HOST scan buffer: pinned host -> device copy of the scan, the
          registration, device -> host copy of the pose/result, all inside the timed region
N > 1 (torchrun): every rank registers its own stream of scans against its own replica of the map — independent
units, no collective on the data path ("scaling": "weak"); value = all ranks' registrations / max-over-ranks time.
`--workload c4` times BASELINE.json configs[3] instead (batched frame pairs, one CTA per pair).
`--impl reference` times the reference's own CPU implementation of the path: its ikd-Tree (compiled unmodified,
oracle/_ref) for the neighbour search + the restated Gauss-Newton loop, on all host cores, same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SCAN, N_MAP, K_NN, ITERS, GATE = 4096, 200000, 5, 20, 2.0
C4_N, C4_ITERS = 2048, 30
POOL = 16   # distinct scans; one C2 step registers all of them against the resident map in ONE call


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_c2(seed=1002):
    from icp4r_loader import pkg
    s = pkg.synth
    sc = s.Scene(seed)
    rng = np.random.default_rng(seed)
    mp = sc.sample(rng, N_MAP)
    scans = []
    for i in range(POOL):
        r = np.random.default_rng(seed * 1000 + i)
        w = sc.sample(r, N_SCAN, radius=40.0)
        scans.append(np.ascontiguousarray(s.apply(np.linalg.inv(s.random_small_se3(r)), w)))
    return mp, scans


NCU_TRAFFIC_BYTES = 7_582_208 + 537_600  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, profiles/r1_c2_reg_iter_kernel_ncu.txt (cold-cache replay)
C5_N, C5_ITERS = 16384, 20


def make_c5(m, seed=1005):
    """C5: dense map, uniform volume density over 400 x 400 x 20 m (6.25 pts/m^3 at 20 M), and a 16,384-pt scan drawn
    from the map's own surfaces (a random subset + 2 cm noise) moved by a small rigid transform."""
    from icp4r_loader import pkg
    s = pkg.synth
    mp = s.dense_map(seed, m)
    rng = np.random.default_rng(seed + 1)
    scans = []
    for i in range(4):
        sel = rng.choice(m, C5_N, replace=False)
        w = mp[sel].copy()
        w[:, :3] += rng.normal(0, 0.02, (C5_N, 3)).astype(np.float32)
        scans.append(np.ascontiguousarray(s.apply(np.linalg.inv(s.random_small_se3(rng, 0.3, 1.0)), w)))
    return mp, scans


def make_c4(n_pairs, seed=1004, with_offsets=False):
    """n_pairs frame pairs of 2,048 points: 64 distinct scenes, each re-used with a different rigid offset of the
    source (generating 65,536 scenes on the host would dominate the run). with_offsets: also return those offsets
    [n_pairs,4,4] (identity for the first 64 pairs)."""
    from icp4r_loader import pkg
    s = pkg.synth
    base = [s.frame_pair(seed + i, C4_N) for i in range(64)]
    rng = np.random.default_rng(seed)
    src = np.empty((n_pairs * C4_N, 4), np.float32)
    tgt = np.empty((n_pairs * C4_N, 4), np.float32)
    D = np.tile(np.eye(4), (n_pairs, 1, 1)) if with_offsets else None
    for p in range(n_pairs):
        a, b, _ = base[p % 64]
        if p >= 64:
            Dp = s.random_small_se3(rng, 0.3, 2.0)
            src[p * C4_N:(p + 1) * C4_N] = s.apply(Dp, a)
            if with_offsets:
                D[p] = Dp
        else:
            src[p * C4_N:(p + 1) * C4_N] = a
        tgt[p * C4_N:(p + 1) * C4_N] = b
    off = (np.arange(n_pairs + 1) * C4_N).astype(np.int32)
    return (src, tgt, off, D) if with_offsets else (src, tgt, off)


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


class quiet_c_stdout:
    """the reference's ikd-Tree printf()s progress lines; keep them off stdout so the JSON line stays alone"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_reference_c2(mp, scans, budget_s, max_regs, threads):
    with quiet_c_stdout():
        return _cpu_reference_c2(mp, scans, budget_s, max_regs, threads)


def _cpu_reference_c2(mp, scans, budget_s, max_regs, threads):
    """The reference's CPU path for this workload: ikd-Tree Build once (resident map), then per registration
    20 x { 4,096 Nearest_Search(k=5, 2.0 m) on `threads` threads + plane fit + 6x6 Gauss-Newton }."""
    import oracle as O
    oo = O.default_opts(residual=O.P2PLANE_KNN, k=K_NN, max_iterations=ITERS, max_corr_dist=GATE)
    if O.have_ref():
        kind, searcher = "reference", O.IkdTree(nthreads=threads)
        searcher.build(mp)
    else:
        kind, searcher = "port", O.BruteSearcher(mp)
        threads = O.num_threads()
    times = []
    t_all = time.perf_counter()
    i = 0
    while i < max_regs and (time.perf_counter() - t_all) < budget_s:
        t0 = time.perf_counter()
        O.register(scans[i % len(scans)], mp, oo, searcher=searcher)
        times.append(time.perf_counter() - t0)
        i += 1
    if hasattr(searcher, "close"):
        searcher.close()
    return kind, threads, times


def run_reference(args, rank, world):
    if rank != 0:
        return
    mp, scans = make_c2()
    threads = os.cpu_count() or 1
    try:
        threads = len(os.sched_getaffinity(0))
    except Exception:
        pass
    # warm-up + timed steps, bounded to a few minutes in total
    # one step = POOL registrations (the scans of one batch of our arm), run one after the other on all host threads;
    # warm-up + timed steps bounded to a few minutes in total
    kind, threads, _ = cpu_reference_c2(mp, scans, 30.0, max(args.warmup, 1), threads)
    kind, threads, times = cpu_reference_c2(mp, scans, 150.0, args.steps * POOL, threads)
    total = float(np.sum(times))
    v = len(times) / total
    nsteps = max(len(times) // POOL, 1)
    line = {
        "impl": "reference", "metric": "registrations/s", "value": v, "unit": "registrations/s", "n_gpus": args.gpus,
        "steps": nsteps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * total / len(times) * POOL, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "nn_queries_per_s": v * ITERS * N_SCAN,
        "config": {"workload": "C2 scan-to-map: 4096-pt scans vs resident 200000-pt map, P2PLANE k=5, 20 iters, gate 2.0 m",
                   "n": N_SCAN, "m": N_MAP, "k": K_NN, "iterations": ITERS, "max_corr_dist": GATE, "batch_scans": POOL,
                   "step": f"{POOL} registrations, one after the other, each on all host threads"},
        "cpu_baseline": {"value": v, "unit": "registrations/s", "cores": threads, "kind": kind,
                         "sample": f"{len(times)} registrations; ikd-Tree Build excluded (resident map); PCL/fast_gicp absent from the image"},
        "e2e": {"value": v, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5", "gicp"])
    ap.add_argument("--frames", type=int, default=2000, help="c3: frames of the odometry sequence (one step = the whole sequence)")
    ap.add_argument("--map-points", type=int, default=20_000_000, help="c5: points of the dense map (whole job)")
    ap.add_argument("--pairs", type=int, default=65536, help="c4: frame pairs per GPU per step")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="c5, N > 1: cross-rank sum inside the iteration kernel over NVLink peer memory, or one NCCL all-reduce per iteration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from icp4r_loader import pkg
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libicp4r_cuda has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h = pkg.Icp4r(local)
    stream = torch.cuda.Stream(device=dev)
    h.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    peak, peak_src = peaks()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """sum of per-step device times (CUDA events on the launching stream); L2 flushed, untimed, before each"""
        evs = []
        with torch.cuda.stream(stream):
            for i in range(steps):
                flush.fill_(i & 0xff)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn(i)
                e1.record(stream)
                evs.append((e0, e1))
        torch.cuda.synchronize(dev)
        return [a.elapsed_time(b) for a, b in evs]

    if args.workload == "c2":
        mp, scans = make_c2(1002 + rank * 0)
        h.map_build(mp)
        o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=K_NN, max_iterations=ITERS, max_corr_dist=GATE)
        off = (np.arange(POOL + 1) * N_SCAN).astype(np.int32)
        cat = np.ascontiguousarray(np.concatenate(scans))
        d_cat = torch.from_numpy(cat).to(dev)
        pin_cat = torch.from_numpy(cat).pin_memory()
        h_cat = pin_cat.numpy()
        d_scans = [torch.from_numpy(s).to(dev) for s in scans]
        units_per_step = POOL
        queries_per_unit = ITERS * N_SCAN
        step_dev = lambda i: h.register_map_batch(d_cat, off, o)
        step_e2e = lambda i: h.register_map_batch(h_cat, off, o)
        h2d, d2h = POOL * N_SCAN * 16, POOL * (16 * 8 + 32)
        cfg = {"workload": "C2 scan-to-map: 4096-pt scans vs resident 200000-pt map, P2PLANE k=5, 20 iters, gate 2.0 m",
               "n": N_SCAN, "m": N_MAP, "k": K_NN, "iterations": ITERS, "max_corr_dist": GATE,
               "batch_scans": POOL, "step": f"{POOL} independent scans registered in one icp4r_register_map_batch call "
               "(one scan alone is latency-bound and leaves most SMs idle; single-scan latency is in latency_ms_single)",
               "l2": "flushed before every timed step (256 MiB fill, untimed)", "replicas": world}
    elif args.workload == "gicp":
        # what radar_odometry.cpp:399-411 runs per frame: fast_gicp's cost (k = 5 covariances, LM, early exit, <= 64 iterations)
        # for one 4,096-pt scan against the resident 200 k-pt map (target normals cached in the handle)
        mp, scans = make_c2(1002)
        h.map_build(mp)
        o = pkg.default_opts(residual=pkg.GICP, k=K_NN, max_iterations=64, early_exit=1, max_corr_dist=0.0)
        d_scans = [torch.from_numpy(s).to(dev) for s in scans]
        pinned = [torch.from_numpy(s).pin_memory() for s in scans]
        h_scans = [p.numpy() for p in pinned]
        units_per_step = 1
        queries_per_unit = 8 * N_SCAN   # ~8 outer iterations until the reference's convergence test fires
        step_dev = lambda i: h.register_map(d_scans[i % POOL], o)
        step_e2e = lambda i: h.register_map(h_scans[i % POOL], o)
        h2d, d2h = N_SCAN * 16, 16 * 8 + 32
        cfg = {"workload": "GICP scan-to-map: 4096-pt scan vs resident 200000-pt map, fast_gicp cost (k=5 covariances, LM), early exit, ungated",
               "n": N_SCAN, "m": N_MAP, "k": K_NN, "max_iterations": 64, "l2": "flushed before every timed step (256 MiB fill, untimed)"}
    elif args.workload == "c3":
        # one step = the whole odometry sequence: Build on frame 0, then per frame register against the growing map,
        # transform with the estimated pose, Add_Points(false). value = frames/s (each frame is one registration).
        seq, _gt = pkg.pipeline.synth_sequence(1003 + rank, args.frames)
        o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=K_NN, max_iterations=ITERS, max_corr_dist=GATE)
        d_seq = [torch.from_numpy(s).to(dev) for s in seq]
        pin_seq = [torch.from_numpy(s).pin_memory() for s in seq]
        h_seq = [p.numpy() for p in pin_seq]
        units_per_step = args.frames
        queries_per_unit = ITERS * int(np.mean([len(s) for s in seq]))
        step_dev = lambda i: pkg.pipeline.run_odometry(h, d_seq, o)
        step_e2e = lambda i: pkg.pipeline.run_odometry(h, h_seq, o)
        h2d, d2h = int(sum(len(s) for s in seq)) * 16, args.frames * (16 * 8 + 32)  # one icp4r_odometry_step per frame: the scan crosses once
        cfg = {"workload": f"C3 odometry sequence: {args.frames} frames (~3000 static pts each), register vs growing map (P2PLANE k=5, 20 iters, gate 2.0 m) + transform + Add_Points(false) per frame (one icp4r_odometry_step call)",
               "frames": args.frames, "k": K_NN, "iterations": ITERS, "max_corr_dist": GATE,
               "final_map_points": int(sum(len(s) for s in seq)), "l2": "flushed before every timed step; the map outgrows L2 during the sequence"}
    elif args.workload == "c5":
        # one map split in spatial slabs along x (halo = gate) across the ranks; every rank gets the same scan; the 29
        # accumulators are summed across ranks every iteration (in-kernel over peer memory, or NCCL). N = 1: the whole map on one GPU.
        mp, scans = make_c5(args.map_points)
        o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=K_NN, max_iterations=C5_ITERS, max_corr_dist=GATE)
        if world > 1:
            bounds = pkg.shard.slab_bounds(mp[:, 0], world)
            mine, lo, hi, _ = pkg.shard.slab_of_rank(mp, rank, world, axis=0, halo=GATE, bounds=bounds)
            uid = [pkg.Icp4r.shard_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            h.shard_init(uid[0], rank, world)
            if args.exchange == "peer":
                hs = [None] * world
                dist.all_gather_object(hs, h.shard_ipc_export())
                h.shard_ipc_import(hs, rank, world)
            h.map_build(torch.from_numpy(mine).to(dev))
            step_dev = lambda i: h.register_sharded(d_scans[i % 4], o, 0, lo, hi)
            step_e2e = lambda i: h.register_sharded(h_scans[i % 4], o, 0, lo, hi)
        else:
            h.map_build(torch.from_numpy(mp).to(dev))
            step_dev = lambda i: h.register_map(d_scans[i % 4], o)
            step_e2e = lambda i: h.register_map(h_scans[i % 4], o)
        d_scans = [torch.from_numpy(s).to(dev) for s in scans]
        pinned = [torch.from_numpy(s).pin_memory() for s in scans]
        h_scans = [p.numpy() for p in pinned]
        units_per_step = 1.0 / world   # ONE registration per step for the whole job (strong scaling)
        queries_per_unit = C5_ITERS * C5_N
        h2d, d2h = C5_N * 16, 16 * 8 + 32
        cfg = {"workload": f"C5 large-map registration: 16384-pt scan vs {args.map_points}-pt dense map in {world} x-slab(s), P2PLANE k=5, 20 iters, gate 2.0 m",
               "n": C5_N, "m": args.map_points, "k": K_NN, "iterations": C5_ITERS, "max_corr_dist": GATE, "slabs": world,
               "collective": ("none" if world == 1 else "29 fp64 sums exchanged inside the iteration kernel over NVLink peer memory" if args.exchange == "peer"
                              else "29-double NCCL all-reduce per iteration"),
               "l2": "flushed before every timed step (256 MiB fill, untimed)"}
    else:
        src, tgt, off = make_c4(args.pairs)
        o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=C4_ITERS)
        d_src, d_tgt, d_off = torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev), torch.from_numpy(off).to(dev)
        outT = torch.zeros((args.pairs, 16), dtype=torch.float64, device=dev)
        outR = torch.zeros((args.pairs, 32), dtype=torch.uint8, device=dev)
        p_src, p_tgt = torch.from_numpy(src).pin_memory(), torch.from_numpy(tgt).pin_memory()
        units_per_step = args.pairs
        queries_per_unit = C4_ITERS * C4_N
        step_dev = lambda i: h.register_batch(d_src, d_off, d_tgt, d_off, o, out=(outT, outR))
        step_e2e = lambda i: h.register_batch(p_src.numpy(), off, p_tgt.numpy(), off, o)
        h2d, d2h = 2 * args.pairs * C4_N * 16 + 2 * (args.pairs + 1) * 4, args.pairs * (128 + 32)
        cfg = {"workload": f"C4 batched registration: {args.pairs} frame pairs per GPU, 2048 pts each, P2P_SVD, 30 iters, ungated",
               "pairs_per_gpu": args.pairs, "n": C4_N, "m": C4_N, "iterations": C4_ITERS,
               "l2": "inputs (%.0f MB per step) exceed L2; flushed anyway" % (2 * args.pairs * C4_N * 16 / 1e6)}

    # ---- warm-up, then the timed device-resident steps ----------------------------------------------------
    timed(step_dev, args.warmup)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = h.launch_count()
    wall0 = time.perf_counter()
    ms = timed(step_dev, args.steps)
    barrier()
    wall = time.perf_counter() - wall0
    launches = h.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    t_dev = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    # ---- end-to-end through the C ABI with host buffers ------------------------------------------------------
    timed(step_e2e, 3)
    barrier()
    ms_e2e = timed(step_e2e, args.steps)
    barrier()
    t_e2e = torch.tensor([sum(ms_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    total_ms, total_e2e_ms = float(t_dev.item()), float(t_e2e.item())
    value = world * args.steps * units_per_step / (total_ms * 1e-3)
    e2e_value = world * args.steps * units_per_step / (total_e2e_ms * 1e-3)

    if rank == 0:
        line = {
            "metric": "registrations/s", "value": value, "unit": "registrations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "c5" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "nn_queries_per_s": value * queries_per_unit,
            "e2e": {"value": e2e_value, "unit": "registrations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": total_e2e_ms / args.steps, "nn_queries_per_s": e2e_value * queries_per_unit},
            "gpu_launches": int(launches), "wall_s_incl_flush": wall, "clocks": clocks,
            "p50_ms": float(np.median(ms)), "p99_ms": float(np.percentile(ms, 99)),
        }
        # ---- roofline of the dominant kernel -------------------------------------------------------------------
        if args.workload == "c2":
            from scipy.spatial import cKDTree
            # single-scan latency through icp4r_register_map (what a sequential odometry loop sees)
            ms1 = timed(lambda i: h.register_map(d_scans[i % POOL], o), max(args.steps // 2, 10))
            line["latency_ms_single"] = float(np.median(ms1))
            line["single_scan_registrations_per_s"] = 1e3 / float(np.mean(ms1))
            # algorithmic bytes of one fused iteration launch over the whole batch (kNN row without the
            # index output the fused kernel never writes): 16*(B*N + M_r) + 232*B, M_r = map points within the gate of
            # >= 1 query of any scan of the batch
            d, _ = cKDTree(cat[:, :3]).query(mp[:, :3], distance_upper_bound=GATE)
            m_r = int(np.isfinite(d).sum())
            alg = 16 * (POOL * N_SCAN + m_r) + 232 * POOL
            # device time of one iteration launch = (t(20 iterations) - t(10 iterations)) / 10, CUDA events, L2 flushed
            o10 = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=K_NN, max_iterations=ITERS // 2, max_corr_dist=GATE)
            t20 = np.mean(timed(step_dev, 10))
            t10 = np.mean(timed(lambda i: h.register_map_batch(d_cat, off, o10), 10))
            k_ms = float(t20 - t10) / (ITERS - ITERS // 2)
            ach = alg / (k_ms * 1e-3) / 1e9
            # traffic: dram__bytes_read.sum + dram__bytes_write.sum of one launch from the ncu --set full capture summarised
            # in profiles/ (ncu flushes caches between replays: the COLD figure; warm launches are served from L2)
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": NCU_TRAFFIC_BYTES,
                                "kernel": "reg_iter_kernel<P2PLANE_KNN,5> (gridDim.y = 16 scans)", "kernel_ms": k_ms, "algorithmic_bytes": alg, "m_r": m_r,
                                "peak_source": peak_src,
                                "note": "working set (3.3 MB map + 1 MB scans) is L2-resident: issue/latency-bound, see DESIGN.md"}
        elif args.workload in ("c5", "c3", "gicp"):
            line["roofline"] = None
        else:
            alg = args.pairs * (16 * 2 * C4_N + 64 + 160)
            k_ms = total_ms / args.steps
            ach = alg / (k_ms * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                "kernel": "reg_batch_kernel<P2P_SVD>", "kernel_ms": k_ms, "algorithmic_bytes": alg,
                                "peak_source": peak_src,
                                "note": "resident design: each pair is read once; the kernel is issue-bound, see DESIGN.md"}
        # ---- CPU baseline beside it (rank 0, N = 1 only, bounded sample) --------------------------------------
        if world == 1 and not args.no_cpu_baseline and args.workload == "c2":
            threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            kind, threads, times = cpu_reference_c2(mp, scans, 15.0, 200, threads)
            v = len(times) / float(np.sum(times))
            line["cpu_baseline"] = {"value": v, "unit": "registrations/s", "cores": threads, "kind": kind,
                                    "sample": f"{len(times)} registrations of the same workload (~15 s); ikd-Tree Build excluded"}
        elif world == 1 and not args.no_cpu_baseline and args.workload == "c4":
            import oracle as O
            oo = O.default_opts(residual=O.P2P_SVD, max_iterations=C4_ITERS)
            t0 = time.perf_counter()
            cnt = 0
            with quiet_c_stdout():
                while time.perf_counter() - t0 < 15.0:
                    a, b = src[cnt * C4_N:(cnt + 1) * C4_N], tgt[cnt * C4_N:(cnt + 1) * C4_N]
                    if O.have_ref():
                        s = O.IkdTree(nthreads=1)
                        s.build(b)
                    else:
                        s = None
                    O.register(a, b, oo, searcher=s)
                    if s is not None:
                        s.close()
                    cnt += 1
            line["cpu_baseline"] = {"value": cnt / (time.perf_counter() - t0), "unit": "registrations/s", "cores": 1,
                                    "kind": "reference" if O.have_ref() else "port",
                                    "sample": f"{cnt} pairs, ikd-Tree Build + 30 x 1-NN + Kabsch each, one thread"}
        elif world == 1 and not args.no_cpu_baseline and args.workload == "c3":
            # the same loop on the host: the reference's ikd-Tree (Build, Add_Points(.., false), Nearest_Search on all
            # threads) + the restated Gauss-Newton loop, on the first frames of the sequence (~20 s)
            import oracle as O
            threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            oo = O.default_opts(residual=O.P2PLANE_KNN, k=K_NN, max_iterations=ITERS, max_corr_dist=GATE)
            with quiet_c_stdout():
                use_ref = O.have_ref()
                T = np.eye(4)
                w0, _ = O.transform(T, seq[0])
                pts = [w0]
                tree = None
                if use_ref:
                    tree = O.IkdTree(nthreads=threads)
                    tree.build(w0)
                t0 = time.perf_counter()
                cnt = 0
                for scan in seq[1:]:
                    for i in range(16):
                        oo.T0[i] = float(T.reshape(16)[i])
                    T, _r, _ = O.register(scan, np.concatenate(pts), oo, searcher=tree)
                    wv, _ = O.transform(T, scan)
                    pts.append(wv)
                    if tree is not None:
                        tree.add_points(wv, False)
                    cnt += 1
                    if time.perf_counter() - t0 > 20.0:
                        break
                dt = time.perf_counter() - t0
                if tree is not None:
                    tree.close()
            line["cpu_baseline"] = {"value": cnt / dt, "unit": "registrations/s", "cores": threads if use_ref else 1,
                                    "kind": "reference" if use_ref else "port",
                                    "sample": f"first {cnt} frames of the sequence (map still small: the reference slows down as the tree grows)"}
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
