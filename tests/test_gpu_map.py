"""GPU parity: map maintenance next to the kNN path — Add_Points with voxel down-sampling and Sector_Search —
against golden vectors produced by the reference's own ikd-Tree and against the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


def run_downsample(handle, base, batches, voxel):
    handle.map_build(base)
    handle.map_set_downsample(voxel)
    rets, valids = [], []
    for b in batches:
        rets.append(handle.map_add_points(b, True))
        valids.append(handle.map_size()[1])
    pts, valid = handle.map_points()
    return rets, valids, pts, valid


@pytest.mark.parametrize("sequential", ["0", "1"])
def test_downsample_golden(handle, O, sequential, monkeypatch):
    """return values, survivor set and kNN over the survivors equal the reference's (both device paths)"""
    monkeypatch.setenv("ICP4R_DS_SEQUENTIAL", sequential)
    g = np.load(os.path.join(G, "downsample.npz"))
    sizes = g["batch_sizes"]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    batches = [g["batches"][offs[i]:offs[i + 1]] for i in range(len(sizes))]
    rets, valids, pts, valid = run_downsample(handle, g["base"], batches, float(g["voxel"]))
    assert rets == list(g["rets"]), (rets, list(g["rets"]))
    assert valids == list(g["validnum"])
    assert (np.nonzero(valid)[0] == g["alive"]).all()
    idx, d2, found = handle.map_knn(g["q"], 5, 0.0)
    assert (found == g["found"]).all() and (bits(d2) == bits(g["d2"])).all()
    same = idx == g["idx"]
    assert same.all() or (bits(d2)[~same] == bits(g["d2"])[~same]).all()  # exact duplicates were inserted on purpose


@pytest.mark.parametrize("voxel", [0.3, 0.5, 0.77, 2.0])
def test_downsample_matches_oracle(handle, O, voxel):
    """voxel sizes that are not powers of two put voxel faces off the float grid (rounding hazards -> exact fallback)"""
    rng = np.random.default_rng(int(voxel * 100))
    base = np.zeros((3000, 4), np.float32)
    base[:, :3] = rng.uniform(-9, 9, (3000, 3)) * np.array([1, 1, 0.2])
    batches = []
    for b in range(3):
        a = np.zeros((1200, 4), np.float32)
        a[:, :3] = rng.uniform(-10, 10, (1200, 3)) * np.array([1, 1, 0.2])
        batches.append(a)
    # points sitting exactly on voxel faces and exact duplicates of existing points
    face = np.zeros((300, 4), np.float32)
    face[:, :3] = (rng.integers(-20, 20, (300, 3)) * np.float32(voxel)).astype(np.float32)
    batches.append(face)
    batches.append(base[:100].copy())
    rets, valids, pts, valid = run_downsample(handle, base, batches, voxel)
    m = O.OracleMap(len(base) + sum(len(b) for b in batches))
    m.add_points(base)
    want = [m.add_points(b, True, voxel) for b in batches]
    assert rets == want
    assert (valid == m.valid[:m.m]).all(), np.nonzero(valid != m.valid[:m.m])[0][:10]
    assert valids[-1] == int(m.valid[:m.m].sum())


def test_downsample_into_empty_and_single_voxel(handle, O):
    rng = np.random.default_rng(3)
    ham = np.zeros((1000, 4), np.float32)
    ham[:, :3] = rng.uniform(0.0, 0.5, (1000, 3)) + np.array([2.0, 2.0, 0.0])
    handle.map_build(ham[:1])
    handle.map_set_downsample(0.5)
    r = handle.map_add_points(ham[1:], True)
    m = O.OracleMap(1000)
    m.add_points(ham[:1])
    assert r == m.add_points(ham[1:], True, 0.5)
    size, nvalid = handle.map_size()
    assert size == 1000 and nvalid == 1 == int(m.valid[:m.m].sum())      # one survivor: the point nearest the voxel centre
    _, valid = handle.map_points()
    assert (valid == m.valid[:m.m]).all()


def test_sector_golden(handle, O):
    g = np.load(os.path.join(G, "sector.npz"))
    handle.map_build(g["pts"])
    for ci, c in enumerate(g["centres"]):
        for hi, hd in enumerate(g["headings"]):
            got = np.sort(handle.map_sector(c, float(g["radius"]), float(hd)))
            want = g[f"s_{ci}_{hi}"]
            assert got.shape == want.shape and (got == want).all(), (ci, hi, len(got), len(want))


def test_sector_matches_oracle_after_downsample(handle, O):
    """deleted points are never returned; centre on a map point (NaN heading) is excluded like in the reference"""
    rng = np.random.default_rng(8)
    pts = np.zeros((5000, 4), np.float32)
    pts[:, :3] = rng.uniform(-60, 60, (5000, 3)) * np.array([1, 1, 0.05])
    handle.map_build(pts[:4000])
    handle.map_set_downsample(2.0)
    handle.map_add_points(pts[4000:], True)
    m = O.OracleMap(5000)
    m.add_points(pts[:4000])
    m.add_points(pts[4000:], True, 2.0)
    for c, r, hd in ((pts[10, :3], 40.0, 30.0), (np.array([1.0, -2.0, 0.0], np.float32), 80.0, -100.0), (np.array([0, 0, 0], np.float32), 15.0, 179.0)):
        got = np.sort(handle.map_sector(c, r, hd))
        want = np.sort(m.sector(c, r, hd))
        assert (got == want).all()


def test_incremental_append_equals_fresh_build(handle, pkg):
    """Add_Points(false) of batches that fit the grid are merged instead of re-sorted: searches over the merged map
    must equal searches over a map built from scratch with the same points (indices, distances, counts — bit for bit),
    through out-of-bounds batches (padded rebuild), density drift (full rebuild) and non-finite points"""
    rng = np.random.default_rng(77)

    def batch(n, cx):
        p = np.zeros((n, 4), np.float32)
        p[:, 0] = rng.uniform(cx - 40, cx + 40, n)
        p[:, 1] = rng.uniform(-40, 40, n)
        p[:, 2] = rng.uniform(-2, 2, n)
        p[:, 3] = rng.uniform(0, 1, n)
        return p

    allp = batch(20000, 0.0)
    handle.map_build(allp)
    fresh = pkg.Icp4r(0)
    try:
        for b in range(36):
            nb = batch(1500, 1.5 * b)            # the window drifts along +x like a moving sensor
            if b == 7:
                nb[::50, 1] = np.nan               # skipped points keep their index
            if b == 20:
                nb = np.concatenate([nb, batch(30000, 1.5 * b)])   # a big batch: density drift
            handle.map_add_points(nb, False)
            allp = np.concatenate([allp, nb])
            if b % 5 == 2 or b in (7, 8, 20, 21, 35):
                q = batch(700, 1.5 * b)
                got = handle.map_knn(q, 5, 2.0)
                fresh.map_build(allp)
                want = fresh.map_knn(q, 5, 2.0)
                assert handle.map_size() == fresh.map_size()
                for a, w in zip(got, want):
                    assert (bits(a) == bits(w)).all(), b
                # far, ungated queries go through the coarse occupancy table, which the merge updates in place
                qf = q[:64].copy()
                qf[:, 0] += 400.0
                for a, w in zip(handle.map_knn(qf, 5, 0.0), fresh.map_knn(qf, 5, 0.0)):
                    assert (bits(a) == bits(w)).all(), b
                o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=4, max_corr_dist=2.0)
                T1, r1, _ = handle.register_map(q, o)
                T2, r2, _ = fresh.register_map(q, o)
                assert np.array_equal(T1, T2) and r1.n_corr == r2.n_corr
        pts, valid = handle.map_points()
        assert (bits(pts) == bits(allp)).all() and valid.all()
    finally:
        fresh.close()


def test_box_radius_delete_golden(handle, O):
    """Box_Search / Radius_Search / Delete_Point_Boxes / Delete_Points / Add_Point_Boxes through the C ABI against the
    reference's ikd-Tree (golden vectors) — index sets, counts, and the k-NN over the survivors after every step"""
    g = np.load(os.path.join(G, "boxops.npz"))
    pts = g["pts"]
    handle.map_build(pts)
    for bi, b in enumerate(g["boxes"]):
        assert (np.sort(handle.map_box_search(b[:3], b[3:])) == g[f"box_{bi}"]).all()
    for ci, c in enumerate(g["centres"]):
        for ri, r in enumerate(g["radii"]):
            assert (np.sort(handle.map_radius_search(c, float(r))) == g[f"rad_{ci}_{ri}"]).all()

    def check(tag):
        _, valid = handle.map_points()
        assert (np.nonzero(valid)[0] == g[f"alive_{tag}"]).all(), tag
        idx, d2, found = handle.map_knn(g["q"], 5, 0.0)
        assert (found == g[f"knn_found_{tag}"]).all() and (bits(d2) == bits(g[f"knn_d2_{tag}"])).all() and (idx == g[f"knn_idx_{tag}"]).all()

    assert handle.map_delete_boxes(g["del_boxes"]) == int(g["del_count"])
    check("after_delete_boxes")
    assert handle.map_delete_points(g["victims"]) == 4
    check("after_delete_points")
    # Add_Point_Boxes: superset of the reference (see tests/test_oracle_golden.py), equal to the oracle
    valid_o = np.ones(len(pts), np.uint8)
    userdel_o = np.zeros(len(pts), np.uint8)
    O.map_delete_boxes(pts, valid_o, userdel_o, g["del_boxes"])
    O.map_delete_points(pts, valid_o, userdel_o, g["victims"])
    want = O.map_add_boxes(pts, valid_o, userdel_o, g["add_boxes"])
    assert handle.map_add_boxes(g["add_boxes"]) == want
    _, valid = handle.map_points()
    assert (valid == valid_o).all() and set(g["alive_after_add_boxes"]) <= set(np.nonzero(valid)[0])
    # points removed by down-sampling never come back; searches on an empty region return nothing
    handle.map_build(pts[:2000])
    handle.map_set_downsample(0.5)
    handle.map_add_points(pts[2000:3000], True)
    _, v0 = handle.map_points()
    assert handle.map_add_boxes(np.array([[-100, -100, -100, 100, 100, 100]], np.float32)) == 0
    _, v1 = handle.map_points()
    assert (v0 == v1).all()
    assert len(handle.map_box_search([500, 500, 500], [501, 501, 501])) == 0


def test_region_search_matches_oracle_large(handle, O):
    rng = np.random.default_rng(9)
    pts = np.zeros((300000, 4), np.float32)
    pts[:, :3] = rng.uniform(-60, 60, (300000, 3)) * np.array([1, 1, 0.1])
    handle.map_build(pts)
    valid = np.ones(len(pts), np.uint8)
    for _ in range(5):
        c = rng.uniform(-50, 50, 3).astype(np.float32) * np.array([1, 1, 0.1], np.float32)
        r = float(rng.uniform(1, 25))
        assert (np.sort(handle.map_radius_search(c, r)) == O.map_radius_search(pts, valid, c, r)).all()
        lo = c - rng.uniform(0.5, 20, 3).astype(np.float32)
        hi = c + rng.uniform(0.5, 20, 3).astype(np.float32)
        assert (np.sort(handle.map_box_search(lo, hi)) == O.map_box_search(pts, valid, lo, hi)).all()


def test_bucket_build_equals_radix_build(pkg, monkeypatch):
    """large maps are built through fixed-capacity buckets (scatter + in-bucket shared-memory sort with every point ranked
    by index inside its cell) instead of the radix sort: the sorted map must be the same — k-NN over both is bit-identical
    and equals the exhaustive search"""
    import torch
    m = 1_200_000
    mp = torch.from_numpy(pkg.synth.dense_map(1005, m)).cuda()
    q = mp[torch.randperm(m, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))[:50000]].clone()
    q[:, :3] += 0.01
    res = {}
    for mode in ("bucket", "radix"):
        monkeypatch.setenv("ICP4R_BUCKET_MIN", "500000" if mode == "bucket" else "-1")
        h = pkg.Icp4r(0)
        n0 = h.launch_count()
        h.map_build(mp)
        launches = h.launch_count() - n0
        idx, d2, found = h.map_knn(q, 5, 2.0)
        h.synchronize()
        res[mode] = (idx.cpu().numpy(), d2.cpu().numpy().view(np.int32), found.cpu().numpy(), launches)
        if mode == "radix":
            bi, bd, bf = h.map_knn_brute(q[:1000], 5, 2.0)
            h.synchronize()
            brute = (bi.cpu().numpy(), bd.cpu().numpy().view(np.int32), bf.cpu().numpy())
        h.close()
    assert res["bucket"][3] < res["radix"][3]          # really two different build paths
    for a, b in zip(res["bucket"][:3], res["radix"][:3]):
        assert np.array_equal(a, b)
    for a, b in zip(res["radix"][:3], brute):
        assert np.array_equal(a[:1000], b)
