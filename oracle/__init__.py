"""CPU oracle bindings — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrappers over ``oracle/liboracle.so`` (plain-C restatement, oracle.c) and
``oracle/_ref/libikd_ref.so`` (the reference's own ikd-Tree compiled unmodified, ref_driver.cpp).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package; the product (``icp-4dradar_b200/``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ACC_LEN = 32

P2P_SVD, P2P_GN, P2PLANE_KNN, P2LINE, GICP, P2PLANE_3PT = range(6)


class OrcOpts(C.Structure):
    _fields_ = [
        ("residual", C.c_int),
        ("k", C.c_int),
        ("max_iterations", C.c_int),
        ("early_exit", C.c_int),
        ("max_corr_dist", C.c_double),
        ("rot_eps", C.c_double),
        ("trans_eps", C.c_double),
        ("mse_abs_eps", C.c_double),
        ("plane_thresh", C.c_double),
        ("T0", C.c_double * 16),
        ("interp_s", C.c_double),
    ]


class OrcResult(C.Structure):
    _fields_ = [
        ("converged", C.c_int),
        ("iterations", C.c_int),
        ("n_corr", C.c_int),
        ("n_fitness", C.c_int),
        ("fitness", C.c_double),
        ("last_cost", C.c_double),
    ]


class OrcCloud(C.Structure):
    _fields_ = [("xyzw", C.c_void_p), ("valid", C.c_void_p), ("m", C.c_int)]


class IkdCbCtx(C.Structure):
    _fields_ = [("tree", C.c_void_p), ("nthreads", C.c_int)]


KNN_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                     C.c_void_p)


def build(force: bool = False) -> None:
    """Compile liboracle.so and (when /root/reference is present) _ref/libikd_ref.so."""
    so = os.path.join(_HERE, "liboracle.so")
    src_newer = (not os.path.exists(so)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(so) for f in ("oracle.c", "oracle.h"))
    ref_so = os.path.join(_HERE, "_ref", "libikd_ref.so")
    need_ref = os.path.exists("/root/reference/third_party/ikd-Tree/ikd_Tree.cpp") and (
        not os.path.exists(ref_so) or os.path.getmtime(os.path.join(_HERE, "ref_driver.cpp")) > os.path.getmtime(ref_so))
    if force or src_newer or need_ref:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


_lib = None
_ref = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        _lib.orc_register.restype = C.c_int
        _lib.orc_accumulate.restype = C.c_int
        _lib.orc_chol6_solve.restype = C.c_int
        _lib.orc_plane_fit.restype = C.c_int
        _lib.orc_map_add_points.restype = C.c_int
        _lib.orc_map_sector.restype = C.c_int
        _lib.orc_voxel_grid.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_icp_p2p_f32.restype = C.c_int
    return _lib


def have_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libikd_ref.so"))


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        build()
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libikd_ref.so"))
        _ref.ikdref_create.restype = C.c_void_p
        _ref.ikdref_create.argtypes = [C.c_float, C.c_float, C.c_float]
        _ref.ikdref_destroy.argtypes = [C.c_void_p]
        for n in ("ikdref_add_points", "ikdref_size", "ikdref_validnum", "ikdref_sector", "ikdref_radius",
                  "ikdref_box", "ikdref_flatten", "ikdref_delete_boxes", "ikdref_hw_threads"):
            getattr(_ref, n).restype = C.c_int
    return _ref


def f4(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4, a.shape
    return a


def default_opts(**kw) -> OrcOpts:
    o = OrcOpts()
    lib().orc_default_opts(C.byref(o))
    for k, v in kw.items():
        if k == "T0":
            v = np.asarray(v, dtype=np.float64).reshape(16)
            for i in range(16):
                o.T0[i] = float(v[i])
        else:
            setattr(o, k, v)
    return o


# ---------------------------------------------------------------------------------------- kNN

def knn(tgt, q, k, max_dist=0.0, valid=None):
    tgt, q = f4(tgt), f4(q)
    nq = q.shape[0]
    idx = np.empty((nq, k), np.int32)
    d2 = np.empty((nq, k), np.float32)
    found = np.empty(nq, np.int32)
    v = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    lib().orc_knn(_p(tgt), _p(v), C.c_int(tgt.shape[0]), _p(q), C.c_int(nq), C.c_int(k), C.c_double(max_dist),
                  _p(idx), _p(d2), _p(found))
    return idx, d2, found


def transform(T, pts):
    pts = f4(pts)
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
    o32 = np.empty_like(pts)
    o64 = np.empty((pts.shape[0], 3), np.float64)
    lib().orc_transform(_p(T), _p(pts), C.c_int(pts.shape[0]), _p(o32), _p(o64))
    return o32, o64


class BruteSearcher:
    """kNN provider for the loops: exhaustive scan over a target cloud."""

    def __init__(self, tgt, valid=None):
        self.tgt = f4(tgt)
        self.valid = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
        self._cloud = OrcCloud(self.tgt.ctypes.data, None if self.valid is None else self.valid.ctypes.data,
                               self.tgt.shape[0])
        self.fn = C.cast(lib().orc_knn_brute_cb, C.c_void_p)
        self.ctx = C.cast(C.pointer(self._cloud), C.c_void_p)


class IkdTree:
    """The reference's KD_TREE<pcl::PointXYZI> (compiled unmodified). Indices ride in `intensity`."""

    def __init__(self, delete_param=0.3, balance_param=0.6, box_length=0.5, nthreads=1):
        # ctor args as at /root/reference/src/radar_odometry.cpp:92
        self.h = C.c_void_p(ref().ikdref_create(delete_param, balance_param, box_length))
        self.n = 0
        self.nthreads = nthreads
        self._cb = IkdCbCtx(self.h, nthreads)
        self.fn = C.cast(ref().ikdref_knn_cb, C.c_void_p)
        self.ctx = C.cast(C.pointer(self._cb), C.c_void_p)

    def close(self):
        if self.h:
            ref().ikdref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, pts):
        pts = f4(pts)
        ref().ikdref_build(self.h, _p(pts), C.c_int(pts.shape[0]), C.c_int(0))
        self.n = pts.shape[0]

    def add_points(self, pts, downsample_on=False):
        pts = f4(pts)
        r = ref().ikdref_add_points(self.h, _p(pts), C.c_int(pts.shape[0]), C.c_int(self.n), C.c_int(int(downsample_on)))
        self.n += pts.shape[0]
        return r

    def set_downsample(self, v):
        ref().ikdref_set_downsample(self.h, C.c_float(v))

    def size(self):
        return ref().ikdref_size(self.h)

    def validnum(self):
        return ref().ikdref_validnum(self.h)

    def knn(self, q, k, max_dist=0.0, nthreads=None):
        q = f4(q)
        nq = q.shape[0]
        idx = np.empty((nq, k), np.int32)
        d2 = np.empty((nq, k), np.float32)
        found = np.empty(nq, np.int32)
        ref().ikdref_knn(self.h, _p(q), C.c_int(nq), C.c_int(k), C.c_double(max_dist), _p(idx), _p(d2), _p(found),
                         C.c_int(self.nthreads if nthreads is None else nthreads))
        return idx, d2, found

    def sector(self, centre, radius, heading):
        c = np.asarray(list(centre) + [0.0], np.float32)
        cap = max(self.n, 1)
        out = np.empty(cap, np.int32)
        n = ref().ikdref_sector(self.h, _p(c), C.c_float(radius), C.c_float(heading), _p(out), C.c_int(cap))
        return out[:min(n, cap)].copy()

    def radius(self, centre, radius):
        c = np.ascontiguousarray(centre, np.float32)
        cap = max(self.n, 1)
        out = np.empty(cap, np.int32)
        ref().ikdref_radius.restype = C.c_int
        n = ref().ikdref_radius(self.h, _p(c), C.c_float(radius), _p(out), C.c_int(cap))
        return out[:n].copy()

    def box(self, bmin, bmax):
        lo, hi = np.ascontiguousarray(bmin, np.float32), np.ascontiguousarray(bmax, np.float32)
        cap = max(self.n, 1)
        out = np.empty(cap, np.int32)
        ref().ikdref_box.restype = C.c_int
        n = ref().ikdref_box(self.h, _p(lo), _p(hi), _p(out), C.c_int(cap))
        return out[:n].copy()

    def delete_boxes(self, boxes):
        b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        ref().ikdref_delete_boxes.restype = C.c_int
        return ref().ikdref_delete_boxes(self.h, _p(b), C.c_int(b.shape[0]))

    def add_boxes(self, boxes):
        b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        ref().ikdref_add_boxes.restype = None
        ref().ikdref_add_boxes(self.h, _p(b), C.c_int(b.shape[0]))

    def delete_points(self, pts):
        pts = f4(pts)
        ref().ikdref_delete_points.restype = None
        ref().ikdref_delete_points(self.h, _p(pts), C.c_int(pts.shape[0]))

    def flatten(self):
        cap = max(self.n, 1)
        out = np.empty(cap, np.int32)
        n = ref().ikdref_flatten(self.h, _p(out), C.c_int(cap))
        return out[:min(n, cap)].copy()


# ---------------------------------------------------------------------------------------- loops

def knn_k_for(residual, k):
    return 1 if residual in (P2P_SVD, P2P_GN, GICP) else (2 if residual == P2LINE else (3 if residual == P2PLANE_3PT else (k if k > 0 else 5)))


def register(src, tgt, opts: OrcOpts, searcher=None, dump=False):
    """Run the restated loop. Returns (T[4,4], OrcResult, dumps or None)."""
    src, tgt = f4(src), f4(tgt)
    s = searcher or BruteSearcher(tgt)
    n, it = src.shape[0], opts.max_iterations
    k = knn_k_for(opts.residual, opts.k)
    T = np.zeros(16, np.float64)
    res = OrcResult()
    dp = np.zeros((it, 16), np.float64) if dump else None
    da = np.zeros((it, ACC_LEN), np.float64) if dump else None
    di = np.full((it, n, k), -1, np.int32) if dump else None
    rc = lib().orc_register(_p(src), C.c_int(n), _p(tgt), C.c_int(tgt.shape[0]), s.fn, s.ctx, C.byref(opts),
                            _p(T), C.byref(res), _p(dp), _p(da), _p(di))
    assert rc == 0
    return T.reshape(4, 4), res, ((dp, da, di) if dump else None)


def accumulate(src, tgt, opts: OrcOpts, T, searcher=None):
    src, tgt = f4(src), f4(tgt)
    s = searcher or BruteSearcher(tgt)
    k = knn_k_for(opts.residual, opts.k)
    T = np.ascontiguousarray(T, np.float64).reshape(16)
    acc = np.zeros(ACC_LEN, np.float64)
    idx = np.full((src.shape[0], k), -1, np.int32)
    used = lib().orc_accumulate(_p(src), C.c_int(src.shape[0]), _p(tgt), C.c_int(tgt.shape[0]), s.fn, s.ctx,
                                C.byref(opts), _p(T), _p(acc), _p(idx))
    return acc, idx, used


def icp_p2p_f32(src, tgt, iters, searcher=None):
    src, tgt = f4(src), f4(tgt)
    s = searcher or BruteSearcher(tgt)
    T = np.zeros(16, np.float32)
    lib().orc_icp_p2p_f32(_p(src), C.c_int(src.shape[0]), _p(tgt), C.c_int(tgt.shape[0]), s.fn, s.ctx,
                          C.c_int(iters), _p(T))
    return T.reshape(4, 4)


# ---------------------------------------------------------------------------------------- GICP

def gicp_normals(pts, k, searcher=None):
    pts = f4(pts)
    s = searcher or BruteSearcher(pts)
    out = np.zeros((pts.shape[0], 3), np.float64)
    lib().orc_gicp_normals(_p(pts), C.c_int(pts.shape[0]), s.fn, s.ctx, C.c_int(k), _p(out))
    return out


def gicp_linearize(src, sn, tgt, tn, opts: OrcOpts, T, searcher=None):
    src, tgt = f4(src), f4(tgt)
    s = searcher or BruteSearcher(tgt)
    T = np.ascontiguousarray(T, np.float64).reshape(16)
    acc = np.zeros(ACC_LEN, np.float64)
    idx = np.full(src.shape[0], -1, np.int32)
    sn, tn = np.ascontiguousarray(sn, np.float64), np.ascontiguousarray(tn, np.float64)
    used = lib().orc_gicp_linearize(_p(src), _p(sn), C.c_int(src.shape[0]), _p(tgt), _p(tn), s.fn, s.ctx, C.byref(opts), _p(T), _p(acc), _p(idx))
    return acc, idx, used


def gicp_register(src, tgt, opts: OrcOpts, searcher=None, normals=None, dump=False):
    """(T, result, (dump_pose, dump_acc) or None). Normals come from each cloud's own opts.k nearest neighbours."""
    src, tgt = f4(src), f4(tgt)
    s = searcher or BruteSearcher(tgt)
    k = opts.k if opts.k > 0 else 20
    sn, tn = normals if normals is not None else (gicp_normals(src, k), gicp_normals(tgt, k, searcher))
    T = np.zeros(16, np.float64)
    res = OrcResult()
    it = opts.max_iterations
    dp = np.zeros((it, 16), np.float64) if dump else None
    da = np.zeros((it, ACC_LEN), np.float64) if dump else None
    lib().orc_gicp_register.restype = C.c_int
    rc = lib().orc_gicp_register(_p(src), _p(sn), C.c_int(src.shape[0]), _p(tgt), _p(tn), C.c_int(tgt.shape[0]), s.fn, s.ctx,
                                 C.byref(opts), _p(T), C.byref(res), _p(dp), _p(da))
    assert rc == 0
    return T.reshape(4, 4), res, ((dp, da) if dump else None)


# ---------------------------------------------------------------------------------------- Doppler filter

class OrcDopplerOut(C.Structure):
    _fields_ = [("A", C.c_double), ("b", C.c_double), ("score", C.c_double), ("v", C.c_double * 3), ("n_static", C.c_int),
                ("best_iteration", C.c_int)]


def doppler_filter(records, iterations=0, seed=1, sigma=0.5, split=0.2):
    rec = np.ascontiguousarray(records, np.float32)
    assert rec.ndim == 2 and rec.shape[1] == 5
    mask = np.zeros(rec.shape[0], np.uint8)
    out = OrcDopplerOut()
    lib().orc_doppler_filter(_p(rec), C.c_int(rec.shape[0]), C.c_int(iterations), C.c_uint64(seed), C.c_double(sigma), C.c_double(split),
                             _p(mask), C.byref(out))
    return mask, out


# ---------------------------------------------------------------------------------------- small algebra

def svd3_rotation(H):
    H = np.ascontiguousarray(H, np.float64).reshape(9)
    R = np.zeros(9, np.float64)
    lib().orc_svd3_rotation(_p(H), _p(R))
    return R.reshape(3, 3)


def chol6_solve(H21, g):
    H21 = np.ascontiguousarray(H21, np.float64)
    g = np.ascontiguousarray(g, np.float64)
    x = np.zeros(6, np.float64)
    rc = lib().orc_chol6_solve(_p(H21), _p(g), _p(x))
    return rc, x


def se3_exp(xi):
    xi = np.ascontiguousarray(xi, np.float64)
    T = np.zeros(16, np.float64)
    lib().orc_se3_exp(_p(xi), _p(T))
    return T.reshape(4, 4)


def gn_update(acc, T):
    """one Gauss-Newton pose update from summed accumulators (H upper triangle [21], g [6]): T <- exp(x^) T with H x = -g"""
    acc = np.asarray(acc, np.float64)
    rc, x = chol6_solve(acc[:21], acc[21:27])
    assert rc == 0
    return se3_exp(x) @ np.asarray(T, np.float64).reshape(4, 4)


def plane_fit(P):
    P = np.ascontiguousarray(P, np.float64)
    n = np.zeros(3, np.float64)
    d = C.c_double(0)
    ok = lib().orc_plane_fit(_p(P), C.c_int(P.shape[0]), _p(n), C.byref(d))
    return ok, n, d.value


def _v(a, n):
    a = np.ascontiguousarray(a, np.float64).reshape(n)
    return a


def res_distance(q, t, p, c):
    r = np.zeros(3)
    lib().orc_res_distance(_p(_v(q, 4)), _p(_v(t, 3)), _p(_v(p, 3)), _p(_v(c, 3)), _p(r))
    return r


def res_plane_norm(q, t, p, n, d):
    r = np.zeros(1)
    lib().orc_res_plane_norm(_p(_v(q, 4)), _p(_v(t, 3)), _p(_v(p, 3)), _p(_v(n, 3)), C.c_double(d), _p(r))
    return r


def res_plane(q, t, p, j, l, m, s):
    r = np.zeros(1)
    lib().orc_res_plane(_p(_v(q, 4)), _p(_v(t, 3)), _p(_v(p, 3)), _p(_v(j, 3)), _p(_v(l, 3)), _p(_v(m, 3)),
                        C.c_double(s), _p(r))
    return r


def res_edge(q, t, p, a, b, s):
    r = np.zeros(3)
    lib().orc_res_edge(_p(_v(q, 4)), _p(_v(t, 3)), _p(_v(p, 3)), _p(_v(a, 3)), _p(_v(b, 3)), C.c_double(s), _p(r))
    return r


# ---------------------------------------------------------------------------------------- map maintenance

class OracleMap:
    """Index-preserving restatement of Build / Add_Points / Sector_Search over a flat array."""

    def __init__(self, cap):
        self.pts = np.zeros((cap, 4), np.float32)
        self.valid = np.zeros(cap, np.uint8)
        self.m = 0

    def add_points(self, pts, downsample_on=False, voxel=0.5):
        pts = f4(pts)
        assert self.m + pts.shape[0] <= self.pts.shape[0]
        m = C.c_int(self.m)
        r = lib().orc_map_add_points(_p(self.pts), _p(self.valid), C.byref(m), _p(pts), C.c_int(pts.shape[0]),
                                     C.c_int(int(downsample_on)), C.c_float(voxel))
        self.m = m.value
        return r

    def knn(self, q, k, max_dist=0.0):
        return knn(self.pts[:self.m], q, k, max_dist, self.valid[:self.m])

    def sector(self, centre, radius, heading):
        c = np.asarray(centre, np.float32)
        out = np.empty(max(self.m, 1), np.int32)
        n = lib().orc_map_sector(_p(self.pts), _p(self.valid), C.c_int(self.m), _p(c), C.c_float(radius),
                                 C.c_float(heading), _p(out), C.c_int(out.shape[0]))
        return out[:n].copy()


def map_box_search(pts, valid, bmin, bmax):
    pts = f4(pts)
    v = np.ascontiguousarray(valid, np.uint8)
    out = np.empty(max(pts.shape[0], 1), np.int32)
    lib().orc_map_box_search.restype = C.c_int
    n = lib().orc_map_box_search(_p(pts), _p(v), C.c_int(pts.shape[0]), _p(np.ascontiguousarray(bmin, np.float32)),
                                 _p(np.ascontiguousarray(bmax, np.float32)), _p(out), C.c_int(out.shape[0]))
    return out[:n].copy()


def map_radius_search(pts, valid, centre, radius):
    pts = f4(pts)
    v = np.ascontiguousarray(valid, np.uint8)
    out = np.empty(max(pts.shape[0], 1), np.int32)
    lib().orc_map_radius_search.restype = C.c_int
    n = lib().orc_map_radius_search(_p(pts), _p(v), C.c_int(pts.shape[0]), _p(np.ascontiguousarray(centre, np.float32)),
                                    C.c_float(radius), _p(out), C.c_int(out.shape[0]))
    return out[:n].copy()


def map_delete_boxes(pts, valid, userdel, boxes):
    """in place on valid / userdel; returns the number of points deleted"""
    b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
    lib().orc_map_delete_boxes.restype = C.c_int
    return lib().orc_map_delete_boxes(_p(f4(pts)), _p(valid), _p(userdel), C.c_int(pts.shape[0]), _p(b), C.c_int(b.shape[0]))


def map_add_boxes(pts, valid, userdel, boxes):
    b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
    lib().orc_map_add_boxes.restype = C.c_int
    return lib().orc_map_add_boxes(_p(f4(pts)), _p(valid), _p(userdel), C.c_int(pts.shape[0]), _p(b), C.c_int(b.shape[0]))


def map_delete_points(pts, valid, userdel, targets):
    t = f4(targets)
    lib().orc_map_delete_points.restype = C.c_int
    return lib().orc_map_delete_points(_p(f4(pts)), _p(valid), _p(userdel), C.c_int(pts.shape[0]), _p(t), C.c_int(t.shape[0]))


def voxel_grid(pts, leaf, valid=None):
    """pcl::VoxelGrid centroid filter restated (oracle.c: orc_voxel_grid). Returns [L,4] float32, ascending leaf index."""
    pts = f4(pts)
    out = np.zeros((max(pts.shape[0], 1), 4), np.float32)
    v = None if valid is None else np.ascontiguousarray(valid, np.uint8)
    n = lib().orc_voxel_grid(_p(pts), _p(v) if v is not None else None, C.c_int(pts.shape[0]), C.c_float(leaf), _p(out), C.c_int(out.shape[0]))
    if n < 0:
        raise ValueError("leaf size too small for the extent of the cloud")
    return out[:n].copy()


def num_threads() -> int:
    return lib().orc_num_threads()
