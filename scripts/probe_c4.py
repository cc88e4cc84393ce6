"""GPU probe: C4 batched-pairs kernel — parity of a few pairs against the oracle, then throughput for the knobs
(ICP4R_RB_THREADS, ICP4R_RB_SLACK, ICP4R_RB_CELL_PTS, ICP4R_NO_HINTS). usage: python scripts/probe_c4.py [pairs]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from icp4r_loader import pkg
import bench

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
src, tgt, off = bench.make_c4(pairs)
dev = torch.device("cuda", 0)
d_src, d_tgt, d_off = torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev), torch.from_numpy(off).to(dev)
o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=30)
STREAM = torch.cuda.Stream(device=dev)


def run(env, check=False, steps=3):
    for k in ("ICP4R_RB_THREADS", "ICP4R_RB_SLACK", "ICP4R_RB_CELL_PTS", "ICP4R_NO_HINTS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    h = pkg.Icp4r(0)
    h.set_stream(STREAM.cuda_stream)
    T, R = h.register_batch(d_src, d_off, d_tgt, d_off, o)
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(STREAM)
        T, R = h.register_batch(d_src, d_off, d_tgt, d_off, o)
        e1.record(STREAM)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"{env}: {ms:.2f} ms / {pairs} pairs = {pairs / ms:.1f} k registrations/s", flush=True)
    h.close()
    return T.cpu().numpy().reshape(-1, 4, 4), R.cpu().numpy().view(pkg.api.RESULT_DTYPE).reshape(-1)


T0, R0 = run({})
Tn, Rn = run({"ICP4R_NO_HINTS": "1"})
print("hints vs no hints: max |dT| = %.3e, fitness max rel %.3e, n_corr equal %s" % (
    np.abs(T0 - Tn).max(), np.max(np.abs(R0["fitness"] - Rn["fitness"]) / Rn["fitness"]), (R0["n_corr"] == Rn["n_corr"]).all()))
import oracle as O
oo = O.default_opts(residual=O.P2P_SVD, max_iterations=30)
worst = 0.0
for p in (0, 1, 63, 64, 65, pairs - 1):
    To, ro, _ = O.register(src[p * 2048:(p + 1) * 2048], tgt[p * 2048:(p + 1) * 2048], oo)
    worst = max(worst, float(np.abs(np.asarray(To).reshape(4, 4) - T0[p]).max()))
    assert abs(ro.fitness - R0["fitness"][p]) <= 1e-9 * max(1.0, ro.fitness), (p, ro.fitness, R0["fitness"][p])
print("vs oracle (6 pairs): max |dT| = %.3e" % worst)
for env in ({"ICP4R_RB_CELL_PTS": "3.0"}, {"ICP4R_RB_CELL_PTS": "4.0"}, {"ICP4R_RB_CELL_PTS": "5.0"}, {"ICP4R_RB_CELL_PTS": "6.0"},
            {"ICP4R_RB_CELL_PTS": "8.0"}, {"ICP4R_RB_CELL_PTS": "12.0"}, {"ICP4R_RB_CELL_PTS": "4.0", "ICP4R_RB_SLACK": "0.1"}):
    run(env)
