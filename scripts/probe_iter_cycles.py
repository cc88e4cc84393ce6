"""GPU probe (needs `make XFLAGS=-DICP4R_KNN_TIMING`): per-query search cycles inside the fused iteration kernel, by
iteration and by the kind of bound the search had (0 none, 1 previous neighbours, 2 the gate)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icp4r_loader import pkg
import bench
mp, scans = bench.make_c2()
h = pkg.Icp4r(0)
h.map_build(mp)
o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=6, max_corr_dist=2.0)
for _ in range(2):
    T, res, dump = h.register_map(scans[0], o, dump=True)
idx = dump[2]            # [it, n, k]
for it in range(6):
    cyc, kind, found3 = idx[it, :, 4], idx[it, :, 3], (idx[it, :, :3] >= 0).sum(1)
    c = np.sort(cyc)
    print(f"iteration {it}: cycles p50 {c[len(c)//2]}  p90 {c[int(len(c)*.9)]}  p99 {c[int(len(c)*.99)]}  max {c[-1]}   kinds {np.bincount(kind, minlength=3)}")
    for kd in range(3):
        m = kind == kd
        if m.any():
            cc = np.sort(cyc[m])
            print(f"    kind {kd}: n {m.sum():5d}  p50 {cc[len(cc)//2]:6d}  p99 {cc[int(len(cc)*.99)]:6d}  max {cc[-1]:6d}")
