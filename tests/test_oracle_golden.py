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
