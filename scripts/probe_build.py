"""GPU probe: stage times of the full grid build (ICP4R_TRACE=1) for a dense uniform map, Add_Points at 3 M, voxel grid.
usage: probe_build.py [points]"""
import os
import sys

import numpy as np

os.environ["ICP4R_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icp4r_loader import pkg

m = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
mp = pkg.synth.dense_map(1005, m)
dev = torch.device("cuda", 0)
d = torch.from_numpy(mp).to(dev)
h = pkg.Icp4r(0)
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h.synchronize()
    e0.record()
    h.map_build(d)
    h.synchronize()
    e1.record()
    torch.cuda.synchronize()
    print(f"map_build({m}) rep {rep}: {e0.elapsed_time(e1):.3f} ms (wall, incl. host syncs)", flush=True)
