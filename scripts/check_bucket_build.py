"""GPU check: the bucket build (large maps) and the radix build give the same sorted map — k-NN over both is bit-identical
and equals the exhaustive search. usage: check_bucket_build.py [points]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from icp4r_loader import pkg

m = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
mp = torch.from_numpy(pkg.synth.dense_map(1005, m)).cuda()
q = mp[torch.randperm(m, device="cuda")[:100000]].clone()
q[:, :3] += 0.01
res = {}
for mode in ("bucket", "radix"):
    os.environ["ICP4R_BUCKET_MIN"] = "1000000" if mode == "bucket" else "-1"
    h = pkg.Icp4r(0)
    h.map_build(mp)
    idx, d2, found = h.map_knn(q, 5, 2.0)
    h.synchronize()  # device outputs are ordered on the handle's stream
    res[mode] = (idx.cpu().numpy(), d2.cpu().numpy().view(np.int32), found.cpu().numpy())
    if mode == "radix":
        bi, bd, bf = h.map_knn_brute(q[:2000], 5, 2.0)
        h.synchronize()
        res["brute"] = (bi.cpu().numpy(), bd.cpu().numpy().view(np.int32), bf.cpu().numpy())
    h.close()
for name, a, b in zip(("idx", "d2", "found"), res["bucket"], res["radix"]):
    bad = np.argwhere(a != b)
    print(name, "mismatches bucket vs radix:", len(bad), bad[:3].tolist())
for name, a, b in zip(("idx", "d2", "found"), res["radix"], res["brute"]):
    print(name, "mismatches radix vs brute (2000 queries):", int((a[:2000] != b).sum()))
for name, a, b in zip(("idx", "d2", "found"), res["bucket"], res["brute"]):
    print(name, "mismatches bucket vs brute (2000 queries):", int((a[:2000] != b).sum()))
ok = all(np.array_equal(a, b) for a, b in zip(res["bucket"], res["radix"]))
print("bucket build == radix build on", m, "points:", ok)
sys.exit(0 if ok else 1)
