#!/usr/bin/env python
"""bench.py — registrations/s and NN queries/s of the registration hot path on B200, with the reference's
CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload all|c1|c2|c2single|c3|c3ref|c4|c5|gicp]

Default (`--workload all`): the headline line is BASELINE.json configs[3] ("C4", the config the metric "at 1/2/4/8
B200" is quoted on and that fits one GPU): ONE fixed job of 65,536 independent frame pairs (2,048 + 2,048 points,
point-to-point ICP, 30 iterations, ungated) split by pair range across the ranks, no collective ("scaling":
"strong"). One step = the whole job.
  value : registrations/s with the clouds already in HBM (device pointers through the C ABI)
  e2e   : the same through the C ABI with pinned HOST clouds: host->device copy of both clouds (4.3 GB per step,
          chunked on a second stream ahead of the kernels), the registrations, device->host copy of poses/results
At N = 1 the same JSON line carries `configs`: one sub-record per other BASELINE config — c1 (single 1,024-pt pair),
c2_single / c2_batch16 (4,096-pt scan vs resident 200 k-pt map, one scan per call / 16 scans per call), c3 (2,000-frame
odometry sequence: register + transform + Add_Points per frame), c3ref (the scan-to-map node's own per-frame flow from raw
radar frames: Doppler filter, Add_Points, Sector_Search, GICP against the sub-map, VoxelGrid), c5 (16,384-pt scan vs
20 M-pt map) and gicp — each with value, e2e, roofline and cpu_baseline.
At N > 1 `configs` carries c5 sharded in N x-slabs with the cross-rank sum inside the iteration kernel over NVLink
peer memory (`--exchange nccl` for the NCCL all-reduce flavour).
`--workload <one>` makes that config the headline line instead (c2: N > 1 = replicas, weak scaling).
`--impl reference` times the reference's own CPU implementation of the headline path: its ikd-Tree (compiled
unmodified, oracle/_ref) for the neighbour search + the restated Kabsch / Gauss-Newton loop, on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SCAN, N_MAP, K_NN, ITERS, GATE = 4096, 200000, 5, 20, 2.0
C1_N, C1_ITERS = 1024, 30
C4_N, C4_ITERS, C4_PAIRS = 2048, 30, 65536
C5_N, C5_ITERS = 16384, 20
POOL = 16   # distinct scans; one c2_batch16 step registers all of them against the resident map in ONE call
FP32_LANES_PER_S = 148 * 128 * 1.965e9   # fp32 lane-instructions per second of one B200 at the maximum SM clock
DIST_EVAL_FP32 = 8                       # sub x3, mul x3, add x2 per squared distance (no FMA: bit-exact with the reference)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_facts(kernel_key):
    """per-launch DRAM traffic and issue statistics of a kernel from this round's ncu --set full capture
    (profiles/r2_ncu_facts.json, written by scripts/ncu_facts.py from the .ncu-rep; None if not captured)"""
    p = os.path.join(ROOT, "profiles", "r2_ncu_facts.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(kernel_key)


def roofline(kernel, kernel_key, alg_bytes, kernel_ms, note, dist_evals=None, extra=None, traffic_scale=1.0):
    """traffic_scale: the ncu capture processed fewer units than one bench launch does (e.g. 2,048 of the 65,536 pairs):
    its DRAM bytes are scaled to the launch `achieved` is quoted on"""
    peak, peak_src = peaks()
    ach = alg_bytes / (kernel_ms * 1e-3) / 1e9
    facts = ncu_facts(kernel_key)
    r = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
         "traffic": (facts["dram_bytes_read"] + facts["dram_bytes_write"]) * traffic_scale if facts else None,
         "kernel": kernel, "kernel_ms": kernel_ms, "algorithmic_bytes": int(alg_bytes), "peak_source": peak_src, "note": note}
    if facts:
        r["ncu"] = {k: facts[k] for k in facts if k not in ("dram_bytes_read", "dram_bytes_write")}
    if dist_evals is not None:
        # compute-side bound for a cache-resident kernel: squared-distance evaluations per second
        # against what the fp32 pipes could issue if they did nothing else
        ceil = FP32_LANES_PER_S / DIST_EVAL_FP32
        rate = dist_evals / (kernel_ms * 1e-3)
        r["compute"] = {"dist_evals_per_launch": int(dist_evals), "dist_evals_per_s": rate, "fp32_ceiling_evals_per_s": ceil,
                        "frac": rate / ceil}
    if extra:
        r.update(extra)
    return r


# --------------------------------------------------------------------------------------------------- generators
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


def make_c5(m, seed=1005):
    """C5: dense map, uniform volume density over 400 x 400 x 20 m (6.25 pts/m^3 at 20 M), and 16,384-pt scans drawn
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


def make_c4(n_pairs, seed=1004, with_offsets=False, first=0):
    """frame pairs [first, first + n_pairs) of the C4 job, 2,048 points per cloud: 64 distinct scenes, each re-used with
    a different rigid offset of the source (generating 65,536 scenes on the host would dominate the run); pair p
    depends on (seed, p) only, so every rank generates exactly its own range. with_offsets: also return those offsets
    [n_pairs,4,4] (identity for the pairs 0..63)."""
    from icp4r_loader import pkg
    s = pkg.synth
    base = [s.frame_pair(seed + i, C4_N) for i in range(64)]
    src = np.empty((n_pairs * C4_N, 4), np.float32)
    tgt = np.empty((n_pairs * C4_N, 4), np.float32)
    D = np.tile(np.eye(4), (n_pairs, 1, 1)) if with_offsets else None
    for j in range(n_pairs):
        p = first + j
        a, b, _ = base[p % 64]
        if p >= 64:
            Dp = s.random_small_se3(np.random.default_rng((seed, p)), 0.3, 2.0)
            src[j * C4_N:(j + 1) * C4_N] = s.apply(Dp, a)
            if with_offsets:
                D[j] = Dp
        else:
            src[j * C4_N:(j + 1) * C4_N] = a
        tgt[j * C4_N:(j + 1) * C4_N] = b
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


# --------------------------------------------------------------------------------------------------- CPU legs
# Every CPU leg runs the reference's own ikd-Tree (compiled unmodified, oracle/_ref: kind "reference") for Build /
# Add_Points / Nearest_Search when it travelled with the repo, else the oracle's exhaustive search (kind "port"),
# plus the restated solve loop of oracle/oracle.c. PCL, FLANN and fast_gicp are absent from the image.

def cpu_c2(mp, scans, budget_s, max_regs, threads, residual=None, k=K_NN, iters=ITERS, gate=GATE, early_exit=0):
    """ikd-Tree Build once (resident map), then per registration iters x { Nearest_Search on `threads` threads + solve }"""
    import oracle as O
    with quiet_c_stdout():
        oo = O.default_opts(residual=O.P2PLANE_KNN if residual is None else residual, k=k, max_iterations=iters, max_corr_dist=gate,
                            early_exit=early_exit)
        if O.have_ref():
            kind, searcher = "reference", O.IkdTree(nthreads=threads)
            searcher.build(mp)
        else:
            kind, searcher = "port", O.BruteSearcher(mp)
            threads = O.num_threads()
        times = []
        tn = None
        t_all = time.perf_counter()
        i = 0
        while i < max_regs and (time.perf_counter() - t_all) < budget_s:
            t0 = time.perf_counter()
            if oo.residual == O.GICP:
                # the map's covariances are computed once, outside the timed calls (our arm caches them in the handle too;
                # fast_gicp itself recomputes them on every setInputTarget): the generous baseline
                if tn is None:
                    tn = O.gicp_normals(mp, k, searcher)
                    t0 = time.perf_counter()
                sc = scans[i % len(scans)]
                O.gicp_register(sc, mp, oo, searcher=searcher, normals=(O.gicp_normals(sc, k), tn))
            else:
                O.register(scans[i % len(scans)], mp, oo, searcher=searcher)
            times.append(time.perf_counter() - t0)
            i += 1
        if hasattr(searcher, "close"):
            searcher.close()
    return kind, threads, times


def cpu_pairs(src, tgt, n_pts, iters, budget_s, threads, max_pairs=1 << 30):
    """frame pairs one after the other on each of `threads` host threads: ikd-Tree Build + iters x 1-NN + Kabsch per pair.
    Returns (kind, pairs done, seconds)."""
    import oracle as O
    oo = O.default_opts(residual=O.P2P_SVD, max_iterations=iters)
    n_avail = min(len(src) // n_pts, max_pairs)
    use_ref = O.have_ref()
    t0 = time.perf_counter()

    def one(p):
        if time.perf_counter() - t0 > budget_s:
            return 0
        a, b = src[p * n_pts:(p + 1) * n_pts], tgt[p * n_pts:(p + 1) * n_pts]
        s = None
        if use_ref:
            s = O.IkdTree(nthreads=1)
            s.build(b)
        O.register(a, b, oo, searcher=s)   # ctypes releases the GIL for the duration of the call
        if s is not None:
            s.close()
        return 1

    with quiet_c_stdout():
        if threads == 1:
            done = 0
            for p in range(n_avail):
                r = one(p)
                if not r:
                    break
                done += r
        else:
            with ThreadPoolExecutor(threads) as ex:
                done = sum(ex.map(one, range(n_avail)))
    return ("reference" if use_ref else "port"), done, time.perf_counter() - t0


def cpu_c3(raw, budget_s, threads, seed=1):
    """the odometry loop on the host from raw radar frames: per frame the Doppler filter, then register against the
    ikd-Tree map (Nearest_Search on all threads + the Gauss-Newton loop), transform, Add_Points(.., false). Every frame is
    timed until a quarter of the budget is gone, then every stride-th frame — with the frames in between inserted untimed
    at the last estimated pose — so that the sample spreads over the whole sequence with the tree grown to each timed
    frame's size instead of stopping while it is small. Returns (kind, frames timed, seconds timed, last frame reached)."""
    import oracle as O
    oo = O.default_opts(residual=O.P2PLANE_KNN, k=K_NN, max_iterations=ITERS, max_corr_dist=GATE)
    use_ref = O.have_ref()
    buf = np.empty((int(sum(len(s) for s in raw)), 4), np.float32)   # the map in insertion order (neighbour indices refer to it)

    def static_of(f):
        mask, _ = O.doppler_filter(raw[f], 0, seed + f)
        return np.ascontiguousarray(raw[f][mask.astype(bool)][:, :4])

    with quiet_c_stdout():
        T = np.eye(4)
        w0, _ = O.transform(T, static_of(0))
        n = len(w0)
        buf[:n] = w0
        tree = None
        if use_ref:
            tree = O.IkdTree(nthreads=threads)
            tree.build(w0)
        t0 = time.perf_counter()
        timed_s, cnt, stride, f = 0.0, 0, 1, 0
        for f in range(1, len(raw)):
            if f % stride == 0:
                ta = time.perf_counter()
                scan = static_of(f)
                for i in range(16):
                    oo.T0[i] = float(T.reshape(16)[i])
                T, _r, _ = O.register(scan, buf[:n], oo, searcher=tree)
                wv, _ = O.transform(T, scan)
                if tree is not None:
                    tree.add_points(wv, False)
                timed_s += time.perf_counter() - ta
                cnt += 1
            else:
                wv, _ = O.transform(T, static_of(f))
                if tree is not None:
                    tree.add_points(wv, False)
            buf[n:n + len(wv)] = wv
            n += len(wv)
            el = time.perf_counter() - t0
            if stride == 1 and el > budget_s * 0.25:
                left = len(raw) - 1 - f
                per = timed_s / max(cnt, 1)
                stride = max(1, int(np.ceil(left * per / max(budget_s * 0.6, 1e-3))))
            if el > budget_s * 1.5:
                break
        if tree is not None:
            tree.close()
    return ("reference" if use_ref else "port"), cnt, timed_s, f


def cpu_c3ref(frames, poses, budget_s, threads, radius=80.0, leaf=0.5, seed=1):
    """the node's own per-frame flow on the host (Doppler filter, transform, ikd-Tree Add_Points, Sector_Search, GICP with
    the reference's ikd-Tree for the searches, voxel grid over the whole map), timed on a few frames spread over the
    sequence; every frame is placed with the same prior poses as in the GPU arm, the frames in between untimed, so that the tree and the map
    have the right size at every timed frame. Returns (kind, frames timed, seconds, list of timed frame numbers)."""
    import oracle as O
    oo = O.default_opts(residual=O.GICP, k=K_NN, max_iterations=64, early_exit=1, max_corr_dist=0.0)
    use_ref = O.have_ref()
    total = sum(len(f) for f in frames)
    buf = np.empty((total, 4), np.float32)
    n = 0
    nf = len(frames)
    picks = sorted(set(int(round(x)) for x in np.linspace(nf * 0.1, nf - 1, 6)))
    timed_s, done, t_all = 0.0, [], time.perf_counter()
    with quiet_c_stdout():
        tree = O.IkdTree(nthreads=threads) if use_ref else None
        om = O.OracleMap(total) if tree is None else None
        for f, rec in enumerate(frames):
            T = np.asarray(poses[f])   # the prior pose the node places the scan with
            timed = f in picks and (time.perf_counter() - t_all) < budget_s
            t0 = time.perf_counter()
            mask, _ = O.doppler_filter(rec, 0, seed + f)
            static = np.ascontiguousarray(rec[mask.astype(bool)][:, :4])
            scan_w, _ = O.transform(T, static)
            buf[n:n + len(scan_w)] = scan_w
            if f == 0:
                tree.build(scan_w) if tree is not None else om.add_points(scan_w, False)
                n += len(scan_w)
                continue
            tree.add_points(scan_w, False) if tree is not None else om.add_points(scan_w, False)
            n += len(scan_w)
            if timed:
                yaw = float(np.degrees(np.arctan2(T[1, 0], T[0, 0])))
                idx = tree.sector(T[:3, 3], radius, yaw) if tree is not None else om.sector(T[:3, 3], radius, yaw)
                sub = np.ascontiguousarray(buf[idx])
                s2 = None
                if use_ref:   # fast_gicp builds a kd-tree over the sub-map on every setInputTarget
                    s2 = O.IkdTree(nthreads=threads)
                    s2.build(sub)
                O.gicp_register(scan_w, sub, oo, searcher=s2)
                if s2 is not None:
                    s2.close()
                O.voxel_grid(buf[:n], leaf)
                timed_s += time.perf_counter() - t0
                done.append(f)
        if tree is not None:
            tree.close()
    return ("reference" if use_ref else "port"), len(done), timed_s, done


def run_reference(args, rank, world):
    """the reference arm: the headline workload on the host cores (rank 0 only)"""
    if rank != 0:
        return
    threads = host_threads()
    wl = "c4" if args.workload == "all" else args.workload
    if wl == "c4":
        # each step = a bounded sample of the 65,536-pair job: `threads` pairs in flight, ~20 s per step
        step_s = float(os.environ.get("ICP4R_REF_STEP_S", "20"))   # seconds of CPU work per step (tests shrink it)
        n_sample = max(threads * 64, 1024) if step_s >= 5 else max(threads * 2, 32)
        src, tgt, off = make_c4(n_sample)
        kind, done, secs = cpu_pairs(src, tgt, C4_N, C4_ITERS, step_s * max(args.warmup, 1), threads, n_sample)   # warm-up
        tot_done, tot_secs = 0, 0.0
        for _ in range(args.steps):
            kind, done, secs = cpu_pairs(src, tgt, C4_N, C4_ITERS, step_s, threads, n_sample)
            tot_done += done
            tot_secs += secs
            if tot_secs > 120.0:
                break
        v = tot_done / tot_secs
        line = {"impl": "reference", "metric": "registrations/s", "value": v, "unit": "registrations/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * C4_PAIRS / v, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "nn_queries_per_s": v * C4_ITERS * C4_N,
                "config": c4_config(args.pairs, 1),
                "cpu_baseline": {"value": v, "unit": "registrations/s", "cores": threads, "kind": kind,
                                 "sample": f"{tot_done} pairs of the job ({n_sample} distinct), one pair per host thread at a time: ikd-Tree Build + "
                                           "30 x 1-NN + Kabsch each; ms_per_step extrapolates to the 65,536-pair job; PCL/FLANN absent from the image"},
                "e2e": {"value": v, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    # C2 (and the other scan-to-map shapes): one registration after the other, Nearest_Search on all host threads
    mp, scans = make_c2()
    cpu_c2(mp, scans, 30.0, max(args.warmup, 1), threads)
    kind, threads, times = cpu_c2(mp, scans, 150.0, args.steps * POOL, threads)
    total = float(np.sum(times))
    v = len(times) / total
    line = {"impl": "reference", "metric": "registrations/s", "value": v, "unit": "registrations/s", "n_gpus": args.gpus,
            "steps": max(len(times) // POOL, 1), "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * total / len(times) * POOL,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "nn_queries_per_s": v * ITERS * N_SCAN, "config": c2_config(1, True),
            "cpu_baseline": {"value": v, "unit": "registrations/s", "cores": threads, "kind": kind,
                             "sample": f"{len(times)} registrations; ikd-Tree Build excluded (resident map); PCL/fast_gicp absent from the image"},
            "e2e": {"value": v, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def c4_config(pairs, world):
    return {"workload": f"C4 batched registration: ONE job of {pairs} independent frame pairs, 2048 + 2048 pts each, P2P_SVD, 30 iterations, "
                        f"ungated, split by pair range across {world} GPU(s), no collective",
            "pairs": pairs, "n": C4_N, "m": C4_N, "iterations": C4_ITERS, "step": "the whole job",
            "l2": "inputs (%.0f MB per step and GPU) exceed L2; flushed anyway" % (2 * pairs / world * C4_N * 16 / 1e6)}


def c2_config(world, batch):
    c = {"workload": "C2 scan-to-map: 4096-pt scans vs resident 200000-pt map, P2PLANE k=5, 20 iters, gate 2.0 m",
         "n": N_SCAN, "m": N_MAP, "k": K_NN, "iterations": ITERS, "max_corr_dist": GATE,
         "l2": "flushed before every timed step (256 MiB fill, untimed)", "replicas": world}
    if batch:
        c.update(batch_scans=POOL, step=f"{POOL} independent scans registered in one icp4r_register_map_batch call")
    else:
        c.update(step="one scan per icp4r_register_map call")
    return c


# --------------------------------------------------------------------------------------------------- GPU legs
class Bench:
    def __init__(self, args, rank, world, local):
        import torch
        import torch.distributed as dist
        from icp4r_loader import pkg
        self.torch, self.dist, self.pkg = torch, dist, pkg
        self.args, self.rank, self.world, self.local = args, rank, world, local
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: libicp4r_cuda has no CPU fallback")
        torch.cuda.set_device(local)
        self.dev = torch.device("cuda", local)
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.h = pkg.Icp4r(local)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.h.set_stream(self.stream.cuda_stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.threads = host_threads()

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, steps):
        """per-step device times (CUDA events on the launching stream); L2 flushed, untimed, before each"""
        torch = self.torch
        evs = []
        with torch.cuda.stream(self.stream):
            for i in range(steps):
                self.flush.fill_(i & 0xff)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self.stream)
                fn(i)
                e1.record(self.stream)
                evs.append((e0, e1))
        torch.cuda.synchronize(self.dev)
        return [a.elapsed_time(b) for a, b in evs]

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def measure(self, step_dev, step_e2e, steps, warmup, units_total, sample_clocks=False):
        """warm-up, K timed device-resident steps, then the same through host buffers. units_total = units of work the
        WHOLE job (all ranks) completes per step. Returns a dict with value / e2e / launches / clocks."""
        self.timed(step_dev, warmup)
        self.barrier()
        sampler = ClockSampler(self.local) if (sample_clocks and self.rank == 0) else None
        l0 = self.h.launch_count()
        wall0 = time.perf_counter()
        ms = self.timed(step_dev, steps)
        self.barrier()
        wall = time.perf_counter() - wall0
        launches = self.h.launch_count() - l0
        clocks = sampler.stop() if sampler else None
        total_ms = self.max_over_ranks(sum(ms))
        self.timed(step_e2e, 3)
        self.barrier()
        ms_e2e = self.timed(step_e2e, steps)
        self.barrier()
        total_e2e = self.max_over_ranks(sum(ms_e2e))
        return {"value": steps * units_total / (total_ms * 1e-3), "ms_per_step": total_ms / steps,
                "e2e_value": steps * units_total / (total_e2e * 1e-3), "e2e_ms_per_step": total_e2e / steps,
                "launches": int(launches), "clocks": clocks, "wall": wall, "p50_ms": float(np.median(ms)), "p99_ms": float(np.percentile(ms, 99)),
                "steps": steps, "warmup": warmup}

    def record(self, m, queries_per_unit, h2d, d2h, cfg):
        r = {"metric": "registrations/s", "value": m["value"], "unit": "registrations/s", "steps": m["steps"], "warmup": m["warmup"],
             "ms_per_step": m["ms_per_step"], "nn_queries_per_s": m["value"] * queries_per_unit, "config": cfg,
             "e2e": {"value": m["e2e_value"], "unit": "registrations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                     "ms_per_step": m["e2e_ms_per_step"], "nn_queries_per_s": m["e2e_value"] * queries_per_unit},
             "gpu_launches": m["launches"], "p50_ms": m["p50_ms"], "p99_ms": m["p99_ms"]}
        return r

    def iter_kernel_ms(self, call):
        """device time of one iteration launch, from the per-launch CUDA events of the handle's profiling mode"""
        self.h.set_profiling(True)
        ts = []
        for i in range(5):
            self.flush.fill_(i)
            call(i)
            p = self.h.last_profile()
            if i >= 2 and len(p) > 1:
                ts.append(float(np.mean(p[:-1])))   # the last entry is the fitness pass
        self.h.set_profiling(False)
        return float(np.mean(ts)) if ts else None

    def stats_of(self, call):
        """(queries searched, squared-distance evaluations, queries settled without a search) of one call"""
        if not hasattr(self.h, "set_stats"):
            return None
        self.h.set_stats(True)
        call(0)
        s = self.h.get_stats()
        self.h.set_stats(False)
        return s

    def m_r(self, map_pts_dev, m, scan):
        """map points within the gate of >= 1 point of `scan` (the M_r of SURVEY.md §8(d)), counted on the device:
        1-NN of every map point in the scan with the gate as max_dist"""
        h2 = self.pkg.Icp4r(self.local)
        h2.map_build(scan)
        total = 0
        step = 4 << 20
        for a in range(0, m, step):
            _i, _d, found = h2.map_knn(map_pts_dev[a:min(a + step, m)], 1, GATE)
            h2.synchronize()   # device outputs are ordered on the handle's own stream
            total += int((found > 0).sum().item())
        h2.close()
        return total

    # ---- C4 -------------------------------------------------------------------------------------------------------
    def run_c4(self, steps, warmup, main):
        torch, pkg, h = self.torch, self.pkg, self.h
        pairs = self.args.pairs
        lo, hi = pkg.shard.pair_range(pairs, self.rank, self.world)
        mine = hi - lo
        src, tgt, off = make_c4(mine, first=lo)
        o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=C4_ITERS)
        d_src, d_tgt, d_off = torch.from_numpy(src).to(self.dev), torch.from_numpy(tgt).to(self.dev), torch.from_numpy(off).to(self.dev)
        outT = torch.zeros((mine, 16), dtype=torch.float64, device=self.dev)
        outR = torch.zeros((mine, 32), dtype=torch.uint8, device=self.dev)
        p_src, p_tgt = torch.from_numpy(src).pin_memory(), torch.from_numpy(tgt).pin_memory()
        step_dev = lambda i: h.register_batch(d_src, d_off, d_tgt, d_off, o, out=(outT, outR))
        step_e2e = lambda i: h.register_batch(p_src.numpy(), off, p_tgt.numpy(), off, o)
        m = self.measure(step_dev, step_e2e, steps, warmup, pairs, sample_clocks=main)
        rec = self.record(m, C4_ITERS * C4_N, 2 * mine * C4_N * 16 + 2 * (mine + 1) * 4, mine * (128 + 32), c4_config(pairs, self.world))
        rec["scaling"] = "strong"
        rec["_m"] = m
        if self.rank == 0:
            st = self.stats_of(lambda i: h.register_batch(d_src[:1024 * C4_N], d_off[:1025], d_tgt[:1024 * C4_N], d_off[:1025], o))
            evals = st[1] / 1024 * mine if st else None
            alg = mine * (16 * 2 * C4_N + 64 + 160)
            extra = None
            if st:
                extra = {"searches_per_registration": st[0] / 1024, "settled_without_search_per_registration": st[2] / 1024,
                         "candidates_per_search": st[1] / max(st[0], 1)}
            rec["roofline"] = roofline("reg_batch_kernel<P2P_SVD,256>", "reg_batch_kernel", alg, m["ms_per_step"],
                                       "resident design: each pair is read from HBM once (65.8 KB) and iterated 31 times in shared memory; "
                                       "the kernel is bound by instruction issue and the per-iteration barrier/solve chain, see DESIGN.md",
                                       dist_evals=evals, extra=extra, traffic_scale=mine / 2048.0)   # the capture ran 2,048 pairs
            if self.world == 1 and not self.args.no_cpu_baseline:
                k1, d1, s1 = cpu_pairs(src, tgt, C4_N, C4_ITERS, 10.0, 1)
                kN, dN, sN = cpu_pairs(src, tgt, C4_N, C4_ITERS, 12.0, self.threads, max_pairs=max(self.threads * 4, 64))
                rec["cpu_baseline"] = {"value": dN / sN, "unit": "registrations/s", "cores": self.threads, "kind": kN,
                                       "sample": f"{dN} pairs of the job, one pair per host thread at a time (ikd-Tree Build + 30 x 1-NN + Kabsch each); "
                                                 "PCL/FLANN absent from the image",
                                       "single_thread": {"value": d1 / s1, "cores": 1, "sample": f"{d1} pairs on one thread"}}
        del d_src, d_tgt, p_src, p_tgt
        return rec

    # ---- C1 -------------------------------------------------------------------------------------------------------
    def run_c1(self, steps, warmup, main=False):
        torch, pkg, h = self.torch, self.pkg, self.h
        pool = [pkg.synth.frame_pair(1001 + i, C1_N) for i in range(8)]
        o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=C1_ITERS)
        d = [(torch.from_numpy(a).to(self.dev), torch.from_numpy(b).to(self.dev)) for a, b, _ in pool]
        pin = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b, _ in pool]
        step_dev = lambda i: h.register(d[i % 8][0], d[i % 8][1], o)
        step_e2e = lambda i: h.register(pin[i % 8][0].numpy(), pin[i % 8][1].numpy(), o)
        m = self.measure(step_dev, step_e2e, steps, warmup, 1, sample_clocks=main)
        rec = self.record(m, C1_ITERS * C1_N, 2 * C1_N * 16, 16 * 8 + 32,
                          {"workload": "C1 single frame pair: 1024 + 1024 pts, P2P_SVD (Kabsch), 30 iterations, ungated (PCL defaults otherwise); "
                                       "one icp4r_register call = ONE kernel launch (resident kernel: grid, 30 iterations, fitness pass in shared memory)",
                           "n": C1_N, "m": C1_N, "iterations": C1_ITERS, "l2": "flushed before every timed step"})
        rec["_m"] = m
        # the whole registration is ONE launch of the resident kernel (a batch of one pair: one CTA, one SM): target grid,
        # 30 iterations, solves and the fitness pass in shared memory. Its time is the step time minus the result copy.
        st = self.stats_of(step_dev)
        rec["us_per_iteration"] = 1e3 * m["ms_per_step"] / (C1_ITERS + 1)
        rec["roofline"] = roofline("reg_batch_kernel<P2P_SVD,256>, one pair (one CTA)", "reg_batch_kernel_c1", 16 * (C1_N + C1_N) + 64 + 160, m["ms_per_step"],
                                   "one CTA on one SM iterates a 32 KB working set in shared memory: a latency chain (31 passes x {search, block "
                                   "reduce, Kabsch}), not a stream; the HBM figure is the two clouds read once",
                                   dist_evals=st[1] if st else None,
                                   extra={"searches_per_registration": st[0], "settled_without_search_per_registration": st[2]} if st else None)
        if not self.args.no_cpu_baseline:
            src = np.concatenate([a for a, _, _ in pool])
            tgt = np.concatenate([b for _, b, _ in pool])
            kind, done, secs = cpu_pairs(np.tile(src, (8, 1)), np.tile(tgt, (8, 1)), C1_N, C1_ITERS, 5.0, 1)
            rec["cpu_baseline"] = {"value": done / secs, "unit": "registrations/s", "cores": 1, "kind": kind,
                                   "sample": f"{done} pairs on one thread (the reference node is single-threaded): ikd-Tree Build + 30 x 1-NN + Kabsch"}
        return rec

    # ---- C2 -------------------------------------------------------------------------------------------------------
    def run_c2(self, steps, warmup, batch, main=False, with_cpu=True):
        torch, pkg, h = self.torch, self.pkg, self.h
        from scipy.spatial import cKDTree
        mp, scans = make_c2(1002)
        h.map_build(mp)
        o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=K_NN, max_iterations=ITERS, max_corr_dist=GATE)
        off = (np.arange(POOL + 1) * N_SCAN).astype(np.int32)
        cat = np.ascontiguousarray(np.concatenate(scans))
        d_cat = torch.from_numpy(cat).to(self.dev)
        pin_cat = torch.from_numpy(cat).pin_memory()
        d_scans = [torch.from_numpy(s).to(self.dev) for s in scans]
        pin = [torch.from_numpy(s).pin_memory() for s in scans]
        if batch:
            step_dev = lambda i: h.register_map_batch(d_cat, off, o)
            step_e2e = lambda i: h.register_map_batch(pin_cat.numpy(), off, o)
            units, h2d, d2h = POOL, POOL * N_SCAN * 16, POOL * (16 * 8 + 32)
        else:
            step_dev = lambda i: h.register_map(d_scans[i % POOL], o)
            step_e2e = lambda i: h.register_map(pin[i % POOL].numpy(), o)
            units, h2d, d2h = 1, N_SCAN * 16, 16 * 8 + 32
        m = self.measure(step_dev, step_e2e, steps, warmup, units * self.world, sample_clocks=main)
        rec = self.record(m, ITERS * N_SCAN, h2d, d2h, c2_config(self.world, batch))
        rec["scaling"] = "weak"
        rec["_m"] = m
        if self.rank == 0:
            if batch:
                d, _ = cKDTree(cat[:, :3]).query(mp[:, :3], distance_upper_bound=GATE)
                o10 = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=K_NN, max_iterations=ITERS // 2, max_corr_dist=GATE)
                t20 = np.mean(self.timed(step_dev, 10))
                t10 = np.mean(self.timed(lambda i: h.register_map_batch(d_cat, off, o10), 10))
                k_ms = float(t20 - t10) / (ITERS - ITERS // 2)
                nsc, key = POOL, "reg_iter_kernel_c2_batch16"
            else:
                d, _ = cKDTree(scans[0][:, :3]).query(mp[:, :3], distance_upper_bound=GATE)
                k_ms = self.iter_kernel_ms(step_dev)
                nsc, key = 1, "reg_iter_kernel_c2_single"
            m_r = int(np.isfinite(d).sum())
            st = self.stats_of(step_dev)
            rec["us_per_iteration"] = 1e3 * k_ms
            rec["roofline"] = roofline(f"reg_iter_kernel<P2PLANE_KNN,5> (gridDim.y = {nsc} scan(s))", key, 16 * (nsc * N_SCAN + m_r) + 232 * nsc, k_ms,
                                       "working set (3.3 MB map + scans) is L2-resident: issue/latency-bound, see DESIGN.md",
                                       dist_evals=(st[1] / (ITERS + 1)) if st else None, extra={"m_r": m_r})
            if with_cpu and self.world == 1 and not self.args.no_cpu_baseline:
                kind, threads, times = cpu_c2(mp, scans, 12.0, 200, self.threads)
                rec["cpu_baseline"] = {"value": len(times) / float(np.sum(times)), "unit": "registrations/s", "cores": threads, "kind": kind,
                                       "sample": f"{len(times)} registrations of the same workload, one after the other, Nearest_Search on all host threads; "
                                                 "ikd-Tree Build excluded (resident map)"}
        return rec

    # ---- GICP -----------------------------------------------------------------------------------------------------
    def run_gicp(self, steps, warmup, main=False):
        torch, pkg, h = self.torch, self.pkg, self.h
        mp, scans = make_c2(1002)
        h.map_build(mp)
        o = pkg.default_opts(residual=pkg.GICP, k=K_NN, max_iterations=64, early_exit=1, max_corr_dist=0.0)
        d_scans = [torch.from_numpy(s).to(self.dev) for s in scans]
        pin = [torch.from_numpy(s).pin_memory() for s in scans]
        its = []

        def step_dev(i):
            T, r, _ = h.register_map(d_scans[i % POOL], o)
            its.append(r.iterations)
        step_e2e = lambda i: h.register_map(pin[i % POOL].numpy(), o)
        m = self.measure(step_dev, step_e2e, steps, warmup, 1, sample_clocks=main)
        n_it = float(np.mean(its[-steps:])) if its else 8.0
        rec = self.record(m, n_it * N_SCAN, N_SCAN * 16, 16 * 8 + 32,
                          {"workload": "GICP scan-to-map (what radar_odometry.cpp:399-411 runs): 4096-pt scan vs resident 200000-pt map, fast_gicp cost "
                                       "(k=5 covariances, LM), early exit, ungated", "n": N_SCAN, "m": N_MAP, "k": K_NN, "max_iterations": 64,
                           "mean_iterations": n_it, "l2": "flushed before every timed step"})
        rec["_m"] = m
        # the registration as a whole against its compulsory traffic: scan + its normals once, then per linearisation the
        # scan (16 B) + per correspondence the target point and its normal (16 + 24 B) + the correspondence record (56 B)
        alg = N_SCAN * (16 + 24) + n_it * N_SCAN * (16 + 16 + 24 + 56)
        rec["roofline"] = roofline("reg_iter_kernel<GICP,1> + gicp_lm_step, whole registration", "gicp_registration", alg, m["ms_per_step"],
                                   "latency-bound: ~8 linearisations x (grid-wide kernel + single-block LM step) on an L2-resident working set")
        if not self.args.no_cpu_baseline:
            import oracle as O
            kind, threads, times = cpu_c2(mp, scans, 10.0, 50, self.threads, residual=O.GICP, iters=64, gate=0.0, early_exit=1)
            rec["cpu_baseline"] = {"value": len(times) / float(np.sum(times)), "unit": "registrations/s", "cores": threads, "kind": kind,
                                   "sample": f"{len(times)} registrations, restated fast_gicp loop with the reference's ikd-Tree for the neighbour searches; the map's "
                                             "covariances computed once outside the timed calls (cached, like our arm; fast_gicp recomputes them per call); "
                                             "fast_gicp itself is absent from the image"}
        return rec

    # ---- C3 -------------------------------------------------------------------------------------------------------
    def run_c3(self, steps, warmup, main=False):
        torch, pkg, h = self.torch, self.pkg, self.h
        frames = self.args.frames
        raw, _gt = pkg.pipeline.synth_radar_sequence(1003 + self.rank, frames, pts_per_frame=4000, extent=400.0, scan_radius=60.0)
        o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=K_NN, max_iterations=ITERS, max_corr_dist=GATE)
        d_raw = [torch.from_numpy(f).to(self.dev) for f in raw]
        pin = [torch.from_numpy(f).pin_memory() for f in raw]
        step_dev = lambda i: pkg.pipeline.run_odometry_raw(h, d_raw, o)
        step_e2e = lambda i: pkg.pipeline.run_odometry_raw(h, pin, o, on_device=lambda r: r.to(self.dev, non_blocking=True))
        m = self.measure(step_dev, step_e2e, steps, warmup, frames, sample_clocks=main)
        npts, _ = h.map_size()
        n_static = npts / frames
        rec = self.record(m, ITERS * int(n_static), int(sum(len(f) for f in raw)) * 20, frames * (16 * 8 + 32),
                          {"workload": f"C3 odometry sequence: {frames} raw radar frames of 4000 records; per frame the Doppler static-point filter (~{int(n_static)} static "
                                       "points), then register vs the growing map (P2PLANE k=5, 20 iters, gate 2.0 m) + transform + Add_Points(false) "
                                       "(icp4r_doppler_static_points + icp4r_odometry_step); one step = the whole sequence",
                           "frames": frames, "k": K_NN, "iterations": ITERS, "max_corr_dist": GATE, "final_map_points": int(npts),
                           "l2": "flushed before every timed step; the map outgrows L2 during the sequence"})
        rec["_m"] = m
        rec["unit_note"] = "registrations/s = frames/s (one registration per frame)"
        rec["ms_per_frame"] = m["ms_per_step"] / frames
        # dominant kernel: the iteration kernel against the FINAL map (state left behind by the last timed sequence)
        last, _ = h.doppler_static_points(d_raw[-1], 0, seed=1 + frames - 1)
        k_ms = self.iter_kernel_ms(lambda i: h.register_map(last, o))
        mpts = h.map_points_dev()
        m_r = self.m_r(mpts, int(npts), last.cpu().numpy())
        st = self.stats_of(lambda i: h.register_map(last, o))
        rec["us_per_iteration"] = 1e3 * k_ms if k_ms else None
        rec["roofline"] = roofline("reg_iter_kernel<P2PLANE_KNN,5> at the final map", "reg_iter_kernel_c3", 16 * (int(last.shape[0]) + m_r) + 232, k_ms,
                                   "per frame: Doppler filter, 20 iteration kernels (latency chain on an L2-resident neighbourhood of the map), incremental Add_Points",
                                   dist_evals=(st[1] / (ITERS + 1)) if st else None, extra={"m_r": m_r})
        if not self.args.no_cpu_baseline:
            kind, cnt, secs, upto = cpu_c3(raw, 20.0, self.threads)
            rec["cpu_baseline"] = {"value": cnt / secs, "unit": "registrations/s", "cores": self.threads if kind == "reference" else 1, "kind": kind,
                                   "sample": f"{cnt} frames timed out of the first {upto}, spread over the sequence with the tree grown to each frame's size "
                                             "(frames in between inserted untimed); per frame Doppler filter + registration + transform + Add_Points"}
        return rec

    # ---- the node's own per-frame flow ------------------------------------------------------------------------------------
    def run_c3ref(self, steps, warmup, main=False):
        torch, pkg, h = self.torch, self.pkg, self.h
        nf = self.args.frames_ref
        # a forward-looking radar: returns within 80 m and +-60 degrees (the sector the node searches its map with)
        frames, gt = pkg.pipeline.synth_radar_sequence(1003 + self.rank, nf, pts_per_frame=4000, extent=400.0, scan_radius=90.0, fov_deg=55.0, max_range=78.0, forward="y")
        o = pkg.default_opts(residual=pkg.GICP, k=K_NN, max_iterations=64, early_exit=1, max_corr_dist=0.0)
        d_frames = [torch.from_numpy(f).to(self.dev) for f in frames]
        pin = [torch.from_numpy(f).pin_memory() for f in frames]
        total = int(sum(len(f) for f in frames))
        vg_out = torch.empty((total, 4), dtype=torch.float32, device=self.dev)
        last = {}

        def step_dev(i):
            last["poses"], last["n_ds"] = pkg.pipeline.run_reference_flow(h, d_frames, o, vg_out=vg_out, priors=gt)
        step_e2e = lambda i: pkg.pipeline.run_reference_flow(h, pin, o, vg_out=vg_out, on_device=lambda r: r.to(self.dev, non_blocking=True), priors=gt)
        m = self.measure(step_dev, step_e2e, steps, warmup, nf, sample_clocks=main)
        n_map, _ = h.map_size()
        rec = self.record(m, 8 * 3000, total * 20, nf * (16 * 8 + 32),
                          {"workload": f"the scan-to-map node's own per-frame flow (radar_odometry.cpp:328,380-429), {nf} raw radar frames of 4000 records: Doppler "
                                       "static-point filter, pointAssociateToMap, Add_Points(false), Sector_Search 80 m, GICP (k=5, LM, early exit) against the sector "
                                       "sub-map, pose chained by left multiplication, VoxelGrid 0.5 m over the whole map; one step = the whole sequence",
                           "frames": nf, "final_map_points": int(n_map), "final_downsampled_points": int(last.get("n_ds", 0)),
                           "l2": "flushed before every timed step"})
        rec["_m"] = m
        rec["unit_note"] = "registrations/s = frames/s (one registration per frame)"
        rec["ms_per_frame"] = m["ms_per_step"] / nf
        # roofline of the step that streams: the voxel grid over the whole map (16 B read per map point, 16 B written per leaf)
        evs = []
        with torch.cuda.stream(self.stream):
            for i in range(5):
                self.flush.fill_(i)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self.stream)
                ds = h.voxel_grid(None, 0.5, out=vg_out)
                e1.record(self.stream)
                evs.append((e0, e1))
        torch.cuda.synchronize(self.dev)
        vg_ms = float(np.median([a.elapsed_time(b) for a, b in evs[1:]]))
        rec["roofline"] = roofline("icp4r_voxel_grid over the final map (key + sort + segmented mean kernels)", "voxel_grid_c3ref",
                                   17 * int(n_map) + 16 * int(ds.shape[0]), vg_ms,
                                   "the per-frame flow is a chain of small launches (filter, sector, sub-map grid + covariances, ~8 GICP linearisations) around "
                                   "two streaming passes over the growing map (sector filter, voxel grid); the voxel grid is the larger one")
        if not self.args.no_cpu_baseline:
            kind, cnt, secs, which = cpu_c3ref(frames, gt, 25.0, self.threads)
            rec["cpu_baseline"] = {"value": cnt / max(secs, 1e-9), "unit": "registrations/s", "cores": self.threads if kind == "reference" else 1, "kind": kind,
                                   "sample": f"frames {which} of the sequence timed whole (filter, transform, Add_Points, Sector_Search, kd-tree over the sub-map, GICP, "
                                             "voxel grid) with the tree grown to that frame; frames in between inserted untimed; fast_gicp / PCL absent from the image"}
        return rec

    # ---- C5 -------------------------------------------------------------------------------------------------------
    def run_c5(self, steps, warmup, main=False):
        torch, pkg, h, dist = self.torch, self.pkg, self.h, self.dist
        world, rank = self.world, self.rank
        mpts = self.args.map_points
        mp, scans = make_c5(mpts)
        o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=K_NN, max_iterations=C5_ITERS, max_corr_dist=GATE)
        d_scans = [torch.from_numpy(s).to(self.dev) for s in scans]
        pin = [torch.from_numpy(s).pin_memory() for s in scans]
        if world > 1:
            bounds = pkg.shard.slab_bounds(mp[:, 0], world)
            mine, lo, hi, _ = pkg.shard.slab_of_rank(mp, rank, world, axis=0, halo=GATE, bounds=bounds)
            uid = [pkg.Icp4r.shard_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            h.shard_init(uid[0], rank, world)
            if self.args.exchange == "peer":
                hs = [None] * world
                dist.all_gather_object(hs, h.shard_ipc_export())
                h.shard_ipc_import(hs, rank, world)
            d_map = torch.from_numpy(mine).to(self.dev)
            h.map_build(d_map)
            step_dev = lambda i: h.register_sharded(d_scans[i % 4], o, 0, lo, hi)
            step_e2e = lambda i: h.register_sharded(pin[i % 4].numpy(), o, 0, lo, hi)
            m_local = len(mine)
        else:
            d_map = torch.from_numpy(mp).to(self.dev)
            h.map_build(d_map)
            step_dev = lambda i: h.register_map(d_scans[i % 4], o)
            step_e2e = lambda i: h.register_map(pin[i % 4].numpy(), o)
            m_local = mpts
        m = self.measure(step_dev, step_e2e, steps, warmup, 1, sample_clocks=main)
        coll = ("none" if world == 1 else "29 fp64 sums exchanged inside the iteration kernel over NVLink peer memory" if self.args.exchange == "peer"
                else "29-double NCCL all-reduce per iteration")
        rec = self.record(m, C5_ITERS * C5_N, C5_N * 16, 16 * 8 + 32,
                          {"workload": f"C5 large-map registration: 16384-pt scan vs {mpts}-pt dense map in {world} x-slab(s), P2PLANE k=5, 20 iters, gate 2.0 m",
                           "n": C5_N, "m": mpts, "k": K_NN, "iterations": C5_ITERS, "max_corr_dist": GATE, "slabs": world, "collective": coll,
                           "l2": "flushed before every timed step (256 MiB fill, untimed)"})
        rec["scaling"] = "strong"
        rec["_m"] = m
        if rank == 0:
            if world == 1:
                k_ms = self.iter_kernel_ms(step_dev)
                st = self.stats_of(step_dev)
            else:
                k_ms, st = m["ms_per_step"] / (C5_ITERS + 1), None
            m_r = self.m_r(d_map, m_local, scans[0])
            rec["us_per_iteration"] = 1e3 * k_ms
            rec["roofline"] = roofline("reg_iter_kernel<P2PLANE_KNN,5>" + ("" if world == 1 else " (rank 0's slab, cross-rank sum fused in)"),
                                       "reg_iter_kernel_c5", 16 * (C5_N + m_r) + 232, k_ms,
                                       "gather-bound: 16,384 queries touch ~1 % of the 320 MB map per iteration; a launch is a latency chain, not a stream",
                                       dist_evals=(st[1] / (C5_ITERS + 1)) if st else None, extra={"m_r": m_r})
            if world == 1:
                t0 = time.perf_counter()
                h.map_build(d_map)
                h.synchronize()
                rec["map_build_ms"] = 1e3 * (time.perf_counter() - t0)
            if world == 1 and not self.args.no_cpu_baseline:
                # bounded sample: the same registrations against the 2 M map points of the central 126 x 126 x 20 m block
                # (same density, 1/10 of the points: a 20 M-point ikd-Tree Build alone takes minutes) with the scan points
                # that fall inside it
                c = np.abs(mp[:, 0]) < 63.25
                c &= np.abs(mp[:, 1]) < 63.25
                sub = np.ascontiguousarray(mp[c])
                sscans = []
                for s in scans:
                    k = (np.abs(s[:, 0]) < 60.0) & (np.abs(s[:, 1]) < 60.0)
                    sscans.append(np.ascontiguousarray(s[k]))
                kind, threads, times = cpu_c2(sub, sscans, 10.0, 50, self.threads, iters=C5_ITERS)
                nq = float(np.mean([len(s) for s in sscans]))
                per_reg = float(np.mean(times)) * (C5_N / nq)   # scaled to the full 16,384-pt scan (cost is per query)
                rec["cpu_baseline"] = {"value": 1.0 / per_reg, "unit": "registrations/s", "cores": threads, "kind": kind,
                                       "sample": f"{len(times)} registrations of the ~{int(nq)} scan points inside the central block against its {len(sub)} map points "
                                                 f"(same density), scaled by 16384/{int(nq)} queries; Build excluded; a 20 M-point ikd-Tree is ~10 % deeper"}
        return rec


    # ---- streaming kernels (HBM-bound by nature: where a roofline fraction means something) ------------------------
    def run_streams(self):
        """Build (KD_TREE::Build, ikd_Tree.cpp:354-365) of the 20 M-point C5 map, VoxelGrid over 20 M points
        (radar_odometry.cpp:426-429) and Add_Points(false) of one 3,000-point scan into a 3 M-point map (radar_odometry.cpp:390):
        device-resident input, CUDA events on the handle's stream around 5 calls each, points/s and the HBM roofline of the
        whole call (all its kernels, host round trips for the grid geometry included)."""
        torch, pkg, h = self.torch, self.pkg, self.h
        peak, peak_src = peaks()

        def timed_call(fn, reps=5, warm=2):
            for _ in range(warm):
                fn()
            ts = []
            for _ in range(reps):
                self.flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                h.synchronize()
                e0.record(self.stream)
                fn()
                h.synchronize()
                e1.record(self.stream)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return float(np.median(ts))

        def rec(name, what, n_pts, ms, alg_bytes, key, note):
            facts = ncu_facts(key)
            ach = alg_bytes / (ms * 1e-3) / 1e9
            return {"metric": "points/s", "value": n_pts / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms, "config": {"workload": what},
                    "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                 "traffic": facts.get("dram_bytes") if facts else None, "kernel": name, "kernel_ms": ms,
                                 "algorithmic_bytes": int(alg_bytes), "peak_source": peak_src, "note": note,
                                 **({"ncu_kernels_us": facts.get("kernels_us")} if facts else {})}}

        out = {}
        m = self.args.map_points
        with torch.cuda.stream(self.stream):
            mp = torch.from_numpy(pkg.synth.dense_map(1005, m)).to(self.dev)
            ms = timed_call(lambda: h.map_build(mp))
            ncell = max(m // 8, 1)
            out["build"] = rec("map build: bbox + bucket scatter + per-bucket shared-memory sort + coarse table", f"KD_TREE::Build of a {m}-point dense map (C5's map), device-resident points",
                               m, ms, 32.0 * m + 4.0 * ncell, "stream_build",
                               "algorithmic = 16 B read + 16 B written per point + the cell table; the build moves every point twice (bucket scatter, in-bucket sort) "
                               "after a bounding-box pass: ~80 B of traffic per point, two host round trips for the grid geometry")
            g = torch.Generator(device=self.dev).manual_seed(7)
            p = torch.rand((m, 4), generator=g, device=self.dev)
            p[:, 0] = (p[:, 0] - 0.5) * 400
            p[:, 1] = (p[:, 1] - 0.5) * 400
            p[:, 2] = (p[:, 2] - 0.5) * 20
            vg_out = torch.empty((m, 4), dtype=torch.float32, device=self.dev)
            leaves = [0]

            def vg():
                leaves[0] = int(h.voxel_grid_into(p, 0.5, vg_out))
            ms = timed_call(vg)
            out["voxel_grid"] = rec("icp4r_voxel_grid: min/max + leaf keys + radix sort + segmented mean", f"pcl::VoxelGrid 0.5 m over {m} points in a 400 x 400 x 20 m volume -> {leaves[0]} leaves",
                                    m, ms, 16.0 * m + 16.0 * leaves[0], "stream_voxel_grid",
                                    "algorithmic = 16 B read per point + 16 B written per leaf; the sort of (leaf key, index) pairs and the per-leaf gather of 16-byte "
                                    "points from 32-byte sectors are what the traffic above that is")
            del p, vg_out, mp
            s_ = pkg.synth
            rng = np.random.default_rng(1003)
            sc = s_.Scene(1003, extent=400.0, n_walls=200)
            h2 = pkg.Icp4r(self.local)
            h2.set_stream(self.stream.cuda_stream)
            m3 = 3_000_000
            h2.map_build(torch.from_numpy(sc.sample(rng, m3)).to(self.dev))
            ts = []
            for f in range(12):
                w = torch.from_numpy(sc.sample(rng, 3000, centre=(10.0 + f, 5.0), radius=80.0)).to(self.dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                h2.synchronize()
                e0.record(self.stream)
                h2.map_add_points(w, False)
                h2.synchronize()
                e1.record(self.stream)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts[2:]))
            out["add_points"] = rec("incremental Add_Points: sort of the new keys + merge into the sorted map + cell-table shift",
                                    f"KD_TREE::Add_Points(3,000 points, false) into a {m3}-point map (C3 per-frame shape)", 3000, ms, 32.0 * m3 + 4.0 * (m3 // 4),
                                    "stream_add_points",
                                    "algorithmic = the merge rewrites the sorted map (16 B read + 16 B written per MAP point) + the cell table; work is O(map), "
                                    "not O(batch): a two-level table would change that (DESIGN.md §9)")
            h2.close()
        if self.world == 1 and not self.args.no_cpu_baseline:
            # the reference's CPU path beside each, on a bounded sample of the same kind of input (points/s, one host thread:
            # ikd-Tree's Build / Add_Points and PCL's VoxelGrid are single-threaded)
            import oracle as O
            import time
            have = O.have_ref()
            nb = 2_000_000
            sample = pkg.synth.dense_map(1005, nb)
            if have:
                t = O.IkdTree()
                t0 = time.perf_counter()
                t.build(sample)
                dt = time.perf_counter() - t0
                out["build"]["cpu_baseline"] = {"value": nb / dt, "unit": "points/s", "cores": 1, "kind": "reference",
                                                "sample": f"KD_TREE::Build of {nb} points of the same dense map (the reference's ikd-Tree compiled unmodified)"}
                w = sc.sample(rng, 3000, centre=(20.0, 5.0), radius=80.0)
                t0 = time.perf_counter()
                t.add_points(w, False)
                dt = time.perf_counter() - t0
                out["add_points"]["cpu_baseline"] = {"value": 3000 / dt, "unit": "points/s", "cores": 1, "kind": "reference",
                                                     "sample": f"KD_TREE::Add_Points(3,000 points, false) into that {nb}-point ikd-Tree"}
                t.close()
            vs = np.random.default_rng(7).random((nb, 4), dtype=np.float32)
            vs[:, 0] = (vs[:, 0] - 0.5) * 400
            vs[:, 1] = (vs[:, 1] - 0.5) * 400
            vs[:, 2] = (vs[:, 2] - 0.5) * 20
            t0 = time.perf_counter()
            O.voxel_grid(vs, 0.5)
            dt = time.perf_counter() - t0
            out["voxel_grid"]["cpu_baseline"] = {"value": nb / dt, "unit": "points/s", "cores": 1, "kind": "port",
                                                 "sample": f"pcl::VoxelGrid restated in C (oracle.c) on {nb} points of the same distribution; PCL itself is absent from the image"}
        return out

    # ---- the adapter boundary --------------------------------------------------------------------------------------
    def run_adapters(self):
        """end-to-end timings through the C++ adapters a maintainer of the reference would compile against (pageable
        pcl::PointCloud<pcl::PointXYZI> clouds, every copy and allocation inside): adapters/bench_adapters.cpp"""
        import subprocess
        exe = os.path.join(ROOT, "icp-4dradar_b200", "adapters", "bench_adapters")
        if not os.path.exists(exe):
            return {"unavailable": "adapters/bench_adapters not built (make -C icp-4dradar_b200/adapters)"}
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, CUDA_VISIBLE_DEVICES=str(self.local)))
        if r.returncode != 0:
            return {"unavailable": f"bench_adapters exited {r.returncode}: {r.stderr[-200:]}"}
        d = json.loads(r.stdout.strip().splitlines()[-1])
        return {"metric": "registrations/s", "value": d["c1_registrations_per_s"], "unit": "registrations/s",
                "config": {"workload": "icp4r::IterativeClosestPoint<PointXYZI>::align on 1,024 + 1,024-point pageable PCL clouds, object constructed per frame "
                                       "(iterative_closest_point.cpp:510-521), wall clock around the call; the other figures time KD_TREE::Build / "
                                       "Nearest_Search / Sector_Search / Add_Points, FastGICPSingleThread::align and VoxelGrid::filter the same way"},
                "e2e": {"value": d["c1_registrations_per_s"], "unit": "registrations/s", "h2d_bytes_per_step": 2 * 1024 * 32 * 2, "d2h_bytes_per_step": 1024 * 16 + 160},
                "wall_ms": d}


def strip(rec):
    rec.pop("_m", None)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "c1", "c2", "c2single", "c3", "c3ref", "c4", "c5", "gicp", "streams"])
    ap.add_argument("--frames-ref", type=int, default=400, help="c3ref: raw radar frames of the node's own per-frame flow")
    ap.add_argument("--frames", type=int, default=2000, help="c3: frames of the odometry sequence (one step = the whole sequence)")
    ap.add_argument("--map-points", type=int, default=20_000_000, help="c5: points of the dense map (whole job)")
    ap.add_argument("--pairs", type=int, default=C4_PAIRS, help="c4: frame pairs of the WHOLE job (split across the GPUs)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="c5, N > 1: cross-rank sum inside the iteration kernel over NVLink peer memory, or one NCCL all-reduce per iteration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="workload all: skip the per-config sub-records")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    B = Bench(args, rank, world, local)
    wl = args.workload
    K, W = args.steps, args.warmup
    configs = None
    if wl in ("all", "c4"):
        rec = B.run_c4(K, W, True)
        if wl == "all" and not args.no_configs:
            configs = {}
            if world == 1:
                configs["c1"] = strip(B.run_c1(max(20 * K, 100), 10))
                configs["c2_single"] = strip(B.run_c2(max(10 * K, 50), 10, batch=False, with_cpu=False))
                configs["c2_batch16"] = strip(B.run_c2(max(10 * K, 50), 10, batch=True))
                if configs["c2_batch16"].get("cpu_baseline"):  # the same registrations on the host: one CPU figure serves both records
                    configs["c2_single"]["cpu_baseline"] = dict(configs["c2_batch16"]["cpu_baseline"])
                configs["gicp"] = strip(B.run_gicp(max(10 * K, 50), 5))
                configs["c3"] = strip(B.run_c3(2, 3))
                configs["c3ref"] = strip(B.run_c3ref(2, 3))
                configs["c5"] = strip(B.run_c5(max(5 * K, 30), 5))
                configs["streams"] = B.run_streams()
                configs["adapters"] = B.run_adapters()
            else:
                configs["c5_sharded"] = strip(B.run_c5(max(5 * K, 30), 5))
    elif wl == "streams":
        if rank == 0:
            print(json.dumps({"streams": B.run_streams(), "adapters": B.run_adapters()}), flush=True)
        B.h.close()
        return
    elif wl == "c1":
        rec = B.run_c1(K, W, main=True)
    elif wl in ("c2", "c2single"):
        rec = B.run_c2(K, W, batch=(wl == "c2"), main=True)
    elif wl == "gicp":
        rec = B.run_gicp(K, W, main=True)
    elif wl == "c3":
        rec = B.run_c3(K, W, main=True)
    elif wl == "c3ref":
        rec = B.run_c3ref(K, W, main=True)
    else:
        rec = B.run_c5(K, W, main=True)
    if rank == 0:
        m = rec.pop("_m", None)
        line = {"metric": rec["metric"], "value": rec["value"], "unit": rec["unit"], "n_gpus": world, "steps": rec["steps"], "warmup": rec["warmup"],
                "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": rec.get("scaling", "weak"), "vs_baseline": None,
                "dtype": "f64", "data": "synthetic"}
        for k, v in rec.items():
            if k not in line:
                line[k] = v
        if m is not None:
            line["clocks"] = m["clocks"]
            line["wall_s_incl_flush"] = m["wall"]
        if configs is not None:
            line["configs"] = configs
        print(json.dumps(line), flush=True)
    B.h.close()
    if world > 1:
        B.dist.destroy_process_group()


if __name__ == "__main__":
    main()
