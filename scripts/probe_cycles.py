"""GPU probe (needs a -DICP4R_KNN_TIMING build): per-query cycle histogram of the stand-alone kNN kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp4r_loader import pkg
import bench
mp, scans = bench.make_c2()
h = pkg.Icp4r(0)
h.map_build(mp)
for k, gate in ((5, 2.0), (1, 0.0)):
    for rep in range(3):
        idx, d2, cyc = h.map_knn(scans[0], k, gate)
    c = np.sort(cyc)
    print(f"k={k} gate={gate}: cycles per query  min {c[0]}  p10 {c[409]}  p50 {c[2048]}  p90 {c[3686]}  p99 {c[4055]}  max {c[-1]}  mean {c.mean():.0f}")
    order = np.argsort(cyc)[-8:]
    print("  slowest queries:", [(int(i), int(cyc[i]), (idx[i] >= 0).sum()) for i in order])
