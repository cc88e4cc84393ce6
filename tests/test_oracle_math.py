"""CPU: the oracle's algebra and residual restatements against independent numpy/scipy implementations."""
import numpy as np
import pytest
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation


def rand_q(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def test_residual_functors_vs_scipy(O):
    """radarFactor.hpp restatements vs scipy quaternion algebra (q = x,y,z,w like Eigen coeffs / para_q)"""
    rng = np.random.default_rng(0)
    for _ in range(50):
        q, t, p, a, b, c = rand_q(rng), rng.normal(size=3), rng.normal(size=3) * 10, rng.normal(size=3), rng.normal(size=3), rng.normal(size=3)
        R = Rotation.from_quat(q)
        pw = R.apply(p) + t
        assert np.allclose(O.res_distance(q, t, p, c), pw - c, atol=1e-12)
        n = a / np.linalg.norm(a)
        assert np.allclose(O.res_plane_norm(q, t, p, n, 0.37), n @ pw + 0.37, atol=1e-12)
        for s in (1.0, 0.3, 0.0):
            # slerp(identity -> q, s): rotation vector scaled by s (shortest path: flip q if w < 0)
            qq = q if q[3] >= 0 else -q
            Rs = Rotation.from_rotvec(Rotation.from_quat(qq).as_rotvec() * s)
            lp = Rs.apply(p) + s * t
            nrm = np.cross(a - b, a - c)
            nrm /= np.linalg.norm(nrm)
            assert np.allclose(O.res_plane(q, t, p, a, b, c, s), (lp - a) @ nrm, atol=1e-9)
            assert np.allclose(O.res_edge(q, t, p, a, b, s), np.cross(lp - a, lp - b) / np.linalg.norm(a - b), atol=1e-9)


def num_jac(f, T, eps=1e-6):
    """d f(exp(xi) T) / d xi at 0 by central differences"""
    from_xi = lambda xi: se3_exp_np(xi) @ T
    J = []
    for i in range(6):
        d = np.zeros(6)
        d[i] = eps
        J.append((f(from_xi(d)) - f(from_xi(-d))) / (2 * eps))
    return np.stack(J, -1)


def se3_exp_np(xi):
    from scipy.linalg import expm
    w, v = xi[:3], xi[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    M[:3, 3] = v
    return expm(M)


def test_se3_exp_chol_svd(O):
    rng = np.random.default_rng(1)
    for sc in (1e-9, 1e-3, 0.5, 3.0):
        xi = rng.normal(size=6) * sc
        assert np.allclose(O.se3_exp(xi), se3_exp_np(xi), atol=1e-12)
    A = rng.normal(size=(20, 6))
    H = A.T @ A
    g = rng.normal(size=6)
    rc, x = O.chol6_solve(H[np.triu_indices(6)], g)
    assert rc == 0 and np.allclose(x, np.linalg.solve(H, -g), rtol=1e-9)
    rc, _ = O.chol6_solve(np.zeros(21), g)
    assert rc != 0
    for _ in range(20):
        P = rng.normal(size=(40, 3)) * rng.uniform(0.1, 30)
        Rt = Rotation.from_rotvec(rng.normal(size=3) * 0.7).as_matrix()
        Q = P @ Rt.T + rng.normal(size=(40, 3)) * 0.01
        Hc = (P - P.mean(0)).T @ (Q - Q.mean(0))
        U, S, Vt = np.linalg.svd(Hc)
        D = np.diag([1, 1, np.sign(np.linalg.det(Vt.T @ U.T))])
        assert np.allclose(O.svd3_rotation(Hc), Vt.T @ D @ U.T, atol=1e-9)
    # reflection case: planar data whose optimal orthogonal map is a reflection
    P = rng.normal(size=(30, 3)) * np.array([1, 1, 0.0])
    Q = P * np.array([1, -1, 1])
    Hc = (P - P.mean(0)).T @ (Q - Q.mean(0))
    R = O.svd3_rotation(Hc)
    assert abs(np.linalg.det(R) - 1) < 1e-9 and np.allclose(R @ R.T, np.eye(3), atol=1e-9)


def test_plane_fit_loam_convention(O):
    rng = np.random.default_rng(2)
    n0 = np.array([0.2, -0.3, 0.93])
    n0 /= np.linalg.norm(n0)
    P = rng.normal(size=(5, 3)) * 0.4 + np.array([12.0, -7.0, 1.0])
    P -= np.outer((P @ n0) - 2.5, n0)  # exactly on the plane n0.x = 2.5
    ok, n, d = O.plane_fit(P)
    assert ok and np.allclose(np.abs(n @ n0), 1, atol=1e-9) and np.allclose(P @ n + d, 0, atol=1e-9)
    x, *_ = np.linalg.lstsq(P, -np.ones(5), rcond=None)  # A n = -1
    assert np.allclose(n, x / np.linalg.norm(x), atol=1e-8) and np.isclose(d, 1 / np.linalg.norm(x), rtol=1e-8)


@pytest.mark.parametrize("kind,k,gate", [("P2P_GN", 1, 0.0), ("P2PLANE_KNN", 5, 2.0), ("P2LINE", 2, 3.0), ("P2PLANE_3PT", 3, 3.0)])
def test_analytic_jacobians_vs_finite_differences(O, pkg, kind, k, gate):
    """J^T r accumulated with the analytic Jacobians == gradient of 0.5*sum r^2 under T <- exp(xi) T with the
    correspondences (and fitted planes/lines) frozen — checked by central differences of the restated functors."""
    kind = getattr(O, kind)
    src, tgt, _ = pkg.synth.frame_pair(9, 300, 2000, extent=25.0)
    T = pkg.synth.se3(0.01, 0.002, -0.003, (0.1, -0.05, 0.02))
    oo = O.default_opts(residual=kind, k=k, max_iterations=1, max_corr_dist=gate)
    acc, idx, used = O.accumulate(src, tgt, oo, T)
    assert used > 20

    def cost(Tm):
        _, pw = O.transform(Tm, src)
        c = 0.0
        for i in range(len(src)):
            nn = idx[i]
            if (nn < 0).any():
                continue
            if kind == O.P2P_GN:
                r = pw[i] - tgt[nn[0], :3].astype(np.float64)
            elif kind == O.P2PLANE_KNN:
                Pn = tgt[nn, :3].astype(np.float64)
                ok, n, d = O.plane_fit(Pn)
                if not ok or (np.abs(Pn @ n + d) > 0.2).any():
                    continue
                r = np.array([n @ pw[i] + d])
            elif kind == O.P2PLANE_3PT:
                j, l, m = (tgt[nn[t], :3].astype(np.float64) for t in range(3))
                nv = np.cross(j - l, j - m)
                r = np.array([(pw[i] - j) @ (nv / np.linalg.norm(nv))])
            else:
                a, b = tgt[nn[0], :3].astype(np.float64), tgt[nn[1], :3].astype(np.float64)
                r = np.cross(pw[i] - a, pw[i] - b) / np.linalg.norm(a - b)
            c += 0.5 * float(r @ r)
        return np.array([c])

    g_num = num_jac(cost, T, 1e-5)[0]
    assert np.isclose(cost(T)[0], 0.5 * acc[27], rtol=1e-9)
    assert np.allclose(g_num, acc[21:27], rtol=2e-5, atol=1e-6 * np.abs(acc[21:27]).max())


def test_icp_loop_vs_scipy(O, pkg):
    """the restated point-to-point loop == an independent numpy/scipy ICP (cKDTree 1-NN + numpy SVD)"""
    src, tgt, _ = pkg.synth.frame_pair(14, 500, 1500, extent=20.0)
    T = np.eye(4)
    tree = cKDTree(tgt[:, :3].astype(np.float64))
    for _ in range(8):
        p = (src[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32).astype(np.float64)
        pd = src[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
        _, j = tree.query(p)
        q = tgt[j, :3].astype(np.float64)
        H = (pd - pd.mean(0)).T @ (q - q.mean(0))
        U, S, Vt = np.linalg.svd(H)
        D = np.diag([1, 1, np.sign(np.linalg.det(Vt.T @ U.T))])
        R = Vt.T @ D @ U.T
        dT = np.eye(4)
        dT[:3, :3] = R
        dT[:3, 3] = q.mean(0) - R @ pd.mean(0)
        T = dT @ T
    To, ro, _ = O.register(src, tgt, O.default_opts(residual=O.P2P_SVD, max_iterations=8))
    assert np.allclose(T, To, atol=1e-9)
    Tg, _, _ = O.register(src, tgt, O.default_opts(residual=O.P2P_GN, max_iterations=8))
    assert np.allclose(Tg, To, atol=1e-4)  # one GN step is a linearisation of the closed form; same fixed point
    # fp32 mirror of PCL's Scalar=float loop stays close to the fp64 loop
    Tf = O.icp_p2p_f32(src, tgt, 8)
    assert np.allclose(Tf, To, atol=5e-3)


def test_voxel_grid_restatement_small_cases():
    """pcl::VoxelGrid restated: hand-checkable cases (leaf order x fastest, float means, NaN skipped)"""
    import oracle as O
    p = np.array([[0.1, 0.1, 0.1, 1.0], [0.4, 0.2, 0.3, 3.0],      # leaf (0,0,0)
                  [0.6, 0.1, 0.1, 5.0],                             # leaf (1,0,0)
                  [0.1, 0.7, 0.1, 7.0],                             # leaf (0,1,0)
                  [-0.2, 0.1, 0.1, 9.0],                            # leaf (-1,0,0): becomes the first column
                  [np.nan, 0.0, 0.0, 11.0]], np.float32)
    out = O.voxel_grid(p, 0.5)
    assert out.shape == (4, 4)
    want = np.array([[-0.2, 0.1, 0.1, 9.0], [0.25, 0.15, 0.2, 2.0], [0.6, 0.1, 0.1, 5.0], [0.1, 0.7, 0.1, 7.0]], np.float32)
    assert np.allclose(out, want, atol=1e-6), out
    # float accumulation in input order: the mean of the two points of leaf (0,0,0) is (fl(0.1+0.4))/2 exactly
    assert out[1, 0] == (np.float32(0.1) + np.float32(0.4)) / np.float32(2)
    assert O.voxel_grid(np.zeros((0, 4), np.float32), 0.5).shape == (0, 4)
    # deleted points are skipped when a validity mask is given
    v = np.array([1, 0, 1, 1, 1, 1], np.uint8)
    out2 = O.voxel_grid(p, 0.5, valid=v)
    assert out2.shape == (4, 4) and out2[1, 3] == 1.0
    # numpy cross-check on a random cloud: same leaf count, same per-leaf means to float tolerance
    rng = np.random.default_rng(0)
    q = np.zeros((5000, 4), np.float32)
    q[:, :3] = rng.uniform(-10, 10, (5000, 3))
    q[:, 3] = rng.uniform(0, 1, 5000)
    o = O.voxel_grid(q, 0.5)
    inv = np.float32(1) / np.float32(0.5)
    ijk = np.floor(q[:, :3] * inv).astype(np.int64)
    mn = ijk.min(0); div = ijk.max(0) - mn + 1
    key = (ijk[:, 0] - mn[0]) + (ijk[:, 1] - mn[1]) * div[0] + (ijk[:, 2] - mn[2]) * div[0] * div[1]
    uk, inv_ix, cnt = np.unique(key, return_inverse=True, return_counts=True)
    assert o.shape[0] == len(uk)
    means = np.zeros((len(uk), 4))
    np.add.at(means, inv_ix, q.astype(np.float64))
    means /= cnt[:, None]
    assert np.allclose(o, means, atol=1e-5)
