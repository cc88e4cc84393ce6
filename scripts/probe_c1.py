"""GPU probe: C1 (one 1,024-point frame pair, P2P_SVD, 30 iterations) through icp4r_register: the resident single-launch path
with 256 / 384 / 512 threads, and the map path (grid in global memory, one launch per iteration)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icp4r_loader import pkg

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
pairs = [pkg.synth.frame_pair(1001 + i, 1024) for i in range(8)]
d = [(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b, _ in pairs]
pin = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b, _ in pairs]
o = pkg.default_opts(residual=pkg.P2P_SVD, max_iterations=30)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, steps=50, warm=5):
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


for env in ({}, {"ICP4R_RB_THREADS": "384"}, {"ICP4R_RB_THREADS": "512"}, {"ICP4R_REGISTER_VIA_MAP": "1"}):
    for k in ("ICP4R_RB_THREADS", "ICP4R_REGISTER_VIA_MAP"):
        os.environ.pop(k, None)
    os.environ.update(env)
    h = pkg.Icp4r(0)
    h.set_stream(stream.cuda_stream)
    t_dev = timed(lambda i: h.register(d[i % 8][0], d[i % 8][1], o))
    t_host = timed(lambda i: h.register(pin[i % 8][0].numpy(), pin[i % 8][1].numpy(), o))
    T, r, _ = h.register(d[0][0], d[0][1], o)
    print(f"{env}: device-resident {t_dev:.3f} ms, host buffers {t_host:.3f} ms | T[0,3] = {T[0, 3]:.12f} fitness {r.fitness:.9f} n_corr {r.n_corr}", flush=True)
    h.close()
