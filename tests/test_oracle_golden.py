"""CPU: the oracle against golden vectors produced by the REFERENCE's own ikd-Tree (tests/golden/make_golden.py)
and, when the compiled reference is present (oracle/_ref), against the reference run live."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


def test_knn_pair_golden(O):
    g = np.load(os.path.join(G, "knn_pair.npz"))
    idx, d2, f = O.knn(g["tgt"], g["src"], 5, 0.0)
    assert (idx == g["idx"]).all() and (bits(d2) == bits(g["d2"])).all() and (f == g["found"]).all()
    idx, d2, f = O.knn(g["tgt"], g["src"], 5, float(g["gate"]))
    assert (idx == g["idx_g"]).all() and (bits(d2) == bits(g["d2_g"])).all() and (f == g["found_g"]).all()
    assert (f < 5).any() and (f > 0).any()  # the gate really bites in this fixture


def test_knn_incremental_golden(O):
    g = np.load(os.path.join(G, "knn_incr.npz"))
    m = O.OracleMap(9000)
    assert m.add_points(g["pts"][:3000]) == 0
    for s in range(3000, 9000, 2000):
        assert m.add_points(g["pts"][s:s + 2000], False) == 0
    assert m.m == int(g["size"]) == int(g["valid"])
    idx, d2, f = m.knn(g["q"], 5, 0.0)
    assert (idx == g["idx"]).all() and (bits(d2) == bits(g["d2"])).all()
    idx, d2, f = m.knn(g["q"], 5, float(g["gate"]))
    assert (idx == g["idx_g"]).all() and (bits(d2) == bits(g["d2_g"])).all() and (f == g["found_g"]).all()


def test_downsample_golden(O):
    """Add_Points(..., true): return values, survivor set and kNN over the survivors match the reference"""
    g = np.load(os.path.join(G, "downsample.npz"))
    base, allb, sizes = g["base"], g["batches"], g["batch_sizes"]
    m = O.OracleMap(len(base) + len(allb))
    m.add_points(base)
    s = 0
    for bi, n in enumerate(sizes):
        r = m.add_points(allb[s:s + n], True, float(g["voxel"]))
        s += n
        assert r == g["rets"][bi], f"batch {bi}"
        assert int(m.valid[:m.m].sum()) == g["validnum"][bi], f"batch {bi}"
    alive = np.nonzero(m.valid[:m.m])[0]
    assert (alive == g["alive"]).all()
    idx, d2, f = m.knn(g["q"], 5, 0.0)
    assert (f == g["found"]).all() and (bits(d2) == bits(g["d2"])).all()
    same = idx == g["idx"]
    # exact duplicates were inserted on purpose: where indices differ the distances must be bit-equal (ties)
    assert same.all() or (bits(d2)[~same] == bits(g["d2"])[~same]).all()


def test_sector_golden(O):
    g = np.load(os.path.join(G, "sector.npz"))
    m = O.OracleMap(len(g["pts"]))
    m.add_points(g["pts"])
    nonempty = 0
    for ci, c in enumerate(g["centres"]):
        for hi, hd in enumerate(g["headings"]):
            got = np.sort(m.sector(c, float(g["radius"]), float(hd)))
            want = g[f"s_{ci}_{hi}"]
            assert got.shape == want.shape and (got == want).all(), (ci, hi)
            nonempty += len(want) > 0
    assert nonempty >= 8


@pytest.mark.skipif(not os.path.exists("/root/reference"), reason="live reference only in the build container")
def test_oracle_vs_live_reference(O, pkg):
    assert O.have_ref()
    rng = np.random.default_rng(123)
    for seed, n, m, k, gate in ((3, 600, 5000, 5, 0.0), (4, 600, 5000, 8, 1.5), (5, 300, 40, 16, 0.0), (6, 100, 3, 5, 0.0)):
        src, tgt, _ = pkg.synth.frame_pair(seed, n, m, extent=float(rng.choice([10, 80])))
        t = O.IkdTree()
        t.build(tgt)
        want = t.knn(src, k, gate)
        got = O.knn(tgt, src, k, gate)
        assert (got[0] == want[0]).all() and (bits(got[1]) == bits(want[1])).all() and (got[2] == want[2]).all()
        t.close()


def test_box_radius_delete_golden(O):
    """Box_Search, Radius_Search, Delete_Point_Boxes, Delete_Points, Add_Point_Boxes restated over flat arrays ==
    the reference's ikd-Tree (index sets, delete count, k-NN over the survivors after every step)"""
    g = np.load(os.path.join(G, "boxops.npz"))
    pts = g["pts"]
    valid = np.ones(len(pts), np.uint8)
    userdel = np.zeros(len(pts), np.uint8)
    for bi, b in enumerate(g["boxes"]):
        assert (O.map_box_search(pts, valid, b[:3], b[3:]) == g[f"box_{bi}"]).all()
    for ci, c in enumerate(g["centres"]):
        for ri, r in enumerate(g["radii"]):
            assert (O.map_radius_search(pts, valid, c, float(r)) == g[f"rad_{ci}_{ri}"]).all()

    def check(tag):
        assert (np.nonzero(valid)[0] == g[f"alive_{tag}"]).all(), tag
        ik, dk, fk = O.knn(pts, g["q"], 5, 0.0, valid=valid)
        assert (fk == g[f"knn_found_{tag}"]).all() and (dk.view(np.int32) == g[f"knn_d2_{tag}"].view(np.int32)).all()
        assert (ik == g[f"knn_idx_{tag}"]).all()

    assert O.map_delete_boxes(pts, valid, userdel, g["del_boxes"]) == int(g["del_count"])
    check("after_delete_boxes")
    assert O.map_delete_points(pts, valid, userdel, g["victims"]) == 4   # the fifth victim is 1 mm off: not the same point
    check("after_delete_points")
    before = valid.copy()
    O.map_add_boxes(pts, valid, userdel, g["add_boxes"])
    # Add_Point_Boxes: the reference can only revive deleted points that no re-balancing rebuild has purged from the
    # tree yet (Rebuild flattens without the deleted points, ikd_Tree.cpp:633-653,1379-1404) — which ones depends on
    # the tree's shape. The restatement revives every point deleted by Delete_*: a superset, and the extra points
    # are exactly deleted points inside the box.
    mine, ref = set(np.nonzero(valid)[0]), set(g["alive_after_add_boxes"])
    assert ref <= mine
    b = g["add_boxes"][0]
    for j in mine - ref:
        assert not before[j] and (b[:3] <= pts[j, :3]).all() and (b[3:] > pts[j, :3]).all()
    assert len(mine) > int(before.sum())
