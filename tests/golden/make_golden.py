"""Generate golden vectors from the REFERENCE ITSELF (run in the build container only).

The reference ships no fixtures or tests, so the pins are produced by running its own
ikd-Tree — compiled unmodified from /root/reference by oracle/Makefile into oracle/_ref/libikd_ref.so —
on seeded inputs, and committing inputs + outputs as small .npz files:

    python tests/golden/make_golden.py

Files written next to this script:
    knn_pair.npz     Build + Nearest_Search, k=5, ungated and gated (2.0 m), 1,024 x 1,024 frame pair
    knn_incr.npz     Build(3000) + 3 x Add_Points(2000, false) + Nearest_Search k=5, as radar_odometry.cpp does
    downsample.npz   Build + Add_Points(..., true) with a 0.5 m voxel: return values and surviving index set
    sector.npz       Sector_Search index sets for several headings (80 m, as radar_odometry.cpp:396)
    boxops.npz       Box_Search / Radius_Search sets, Delete_Point_Boxes counts, Delete_Points, Add_Point_Boxes: the
                     surviving index set and a k-NN over it after every step
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle as O  # noqa: E402
from icp4r_loader import pkg  # noqa: E402

synth = pkg.synth


def main():
    assert os.path.exists("/root/reference/third_party/ikd-Tree/ikd_Tree.cpp"), "needs the reference checkout"
    O.build(force=True)

    # --- knn_pair
    src, tgt, _ = synth.frame_pair(1001, 1024)
    t = O.IkdTree()
    t.build(tgt)
    i0, d0, f0 = t.knn(src, 5, 0.0)
    i1, d1, f1 = t.knn(src, 5, 2.0)
    np.savez_compressed(os.path.join(HERE, "knn_pair.npz"), src=src, tgt=tgt, idx=i0, d2=d0, found=f0, idx_g=i1, d2_g=d1,
                        found_g=f1, gate=2.0)
    t.close()

    # --- knn_incr
    _, pts, _ = synth.frame_pair(31, 16, 9000)
    q, _, _ = synth.frame_pair(32, 512, 16)
    t = O.IkdTree()
    t.build(pts[:3000])
    for s in range(3000, 9000, 2000):
        assert t.add_points(pts[s:s + 2000], False) == 0
    i0, d0, f0 = t.knn(q, 5, 0.0)
    i1, d1, f1 = t.knn(q, 5, 3.0)
    np.savez_compressed(os.path.join(HERE, "knn_incr.npz"), pts=pts, q=q, idx=i0, d2=d0, found=f0, idx_g=i1, d2_g=d1, found_g=f1,
                        gate=3.0, size=t.size(), valid=t.validnum())
    t.close()

    # --- downsample
    rng = np.random.default_rng(41)
    base = np.zeros((1500, 4), np.float32)
    base[:, :3] = rng.uniform(-6, 6, (1500, 3)) * np.array([1, 1, 0.25])
    batches = []
    for b in range(4):
        a = np.zeros((800, 4), np.float32)
        a[:, :3] = rng.uniform(-7, 7, (800, 3)) * np.array([1, 1, 0.25])
        batches.append(a)
    # one voxel hammered by 1,000 points (SURVEY.md BASELINE probe) and exact duplicates of existing points
    ham = np.zeros((1000, 4), np.float32)
    ham[:, :3] = rng.uniform(0.0, 0.5, (1000, 3)) + np.array([2.0, 2.0, 0.0])
    batches.append(ham.astype(np.float32))
    batches.append(base[:200].copy())
    t = O.IkdTree()
    t.build(base)
    t.set_downsample(0.5)
    rets, sizes = [], []
    for a in batches:
        rets.append(t.add_points(a, True))
        sizes.append(t.validnum())
    alive = np.sort(t.flatten())
    qd = np.zeros((256, 4), np.float32)
    qd[:, :3] = rng.uniform(-7, 7, (256, 3)) * np.array([1, 1, 0.25])
    ik, dk, fk = t.knn(qd, 5, 0.0)
    np.savez_compressed(os.path.join(HERE, "downsample.npz"), base=base, batches=np.concatenate(batches),
                        batch_sizes=np.array([len(a) for a in batches]), rets=np.array(rets), validnum=np.array(sizes),
                        alive=alive, voxel=0.5, q=qd, idx=ik, d2=dk, found=fk)
    t.close()

    # --- sector
    _, pts, _ = synth.frame_pair(51, 16, 6000)
    t = O.IkdTree()
    t.build(pts)
    centres = np.array([[0, 0, 0], [10.5, -20.25, 0.5], [-70, 60, -1]], np.float32)
    headings = np.array([0.0, 45.0, 170.0, -120.0], np.float32)
    out = {}
    for ci, c in enumerate(centres):
        for hi, hd in enumerate(headings):
            out[f"s_{ci}_{hi}"] = np.sort(t.sector(c, 80.0, float(hd)))
    np.savez_compressed(os.path.join(HERE, "sector.npz"), pts=pts, centres=centres, headings=headings, radius=80.0, **out)
    t.close()
    # --- box / radius search and the deletes (ikd_Tree.h:243-249)
    rng = np.random.default_rng(61)
    pts = np.zeros((5000, 4), np.float32)
    pts[:, :3] = rng.uniform(-20, 20, (5000, 3)) * np.array([1, 1, 0.15])
    t = O.IkdTree()
    t.build(pts)
    boxes = np.array([[-5, -5, -1, 5, 5, 1], [3, 3, -3, 12, 9, 3], [-19.5, 10, -0.5, -12, 19, 0.5], [100, 100, 100, 101, 101, 101]], np.float32)
    centres = np.array([[0, 0, 0], [10.5, -12.25, 0.5], [-18, 18, -1]], np.float32)
    radii = np.array([3.0, 7.5, 0.8], np.float32)
    out = {}
    for bi, b in enumerate(boxes):
        out[f"box_{bi}"] = np.sort(t.box(b[:3], b[3:]))
    for ci, c in enumerate(centres):
        for ri, r in enumerate(radii):
            out[f"rad_{ci}_{ri}"] = np.sort(t.radius(c, float(r)))
    steps = []
    q = np.zeros((400, 4), np.float32)
    q[:, :3] = rng.uniform(-20, 20, (400, 3)) * np.array([1, 1, 0.15])

    def snap(tag):
        alive = np.sort(t.flatten())
        ik, dk, fk = t.knn(q, 5, 0.0)
        out[f"alive_{tag}"] = alive
        out[f"knn_idx_{tag}"], out[f"knn_d2_{tag}"], out[f"knn_found_{tag}"] = ik, dk, fk
        steps.append(tag)

    del_boxes = np.array([[-4, -4, -2, 4, 4, 2], [8, -15, -3, 15, -8, 3]], np.float32)
    out["del_count"] = np.int32(t.delete_boxes(del_boxes))
    snap("after_delete_boxes")
    victims = pts[np.array([11, 222, 3333, 4444, 4999])].copy()
    victims[2, 0] += np.float32(5e-7)       # inside EPSS: still the same point
    victims[4, 1] += np.float32(1e-3)       # outside EPSS: nothing to delete
    t.delete_points(victims)
    snap("after_delete_points")
    add_boxes = np.array([[-2, -2, -2, 2, 2, 2]], np.float32)
    t.add_boxes(add_boxes)
    snap("after_add_boxes")
    np.savez_compressed(os.path.join(HERE, "boxops.npz"), pts=pts, boxes=boxes, centres=centres, radii=radii, q=q, del_boxes=del_boxes,
                        victims=victims, add_boxes=add_boxes, **out)
    t.close()
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
