"""ctypes bindings over include/icp4r.h — the same boundary a reference maintainer would bind.

No compute happens here and nothing falls back to the CPU: if ``libicp4r_cuda.so`` is missing the loader
raises, and if there is no CUDA device ``icp4r_create`` fails with ICP4R_ERR_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

HOST, DEVICE = 0, 1
P2P_SVD, P2P_GN, P2PLANE_KNN, P2LINE, GICP, P2PLANE_3PT = range(6)
ACC_LEN = 32
MAX_K = 16

EXPORTS = [
    "icp4r_create", "icp4r_destroy", "icp4r_last_error", "icp4r_version", "icp4r_default_opts", "icp4r_set_stream", "icp4r_set_point_layout",
    "icp4r_synchronize", "icp4r_launch_count", "icp4r_set_profiling", "icp4r_last_profile", "icp4r_set_stats", "icp4r_get_stats", "icp4r_map_build", "icp4r_map_set_downsample", "icp4r_map_add_points",
    "icp4r_map_size", "icp4r_map_range", "icp4r_map_knn", "icp4r_map_knn_brute", "icp4r_map_sector", "icp4r_map_points",
    "icp4r_register", "icp4r_register_map", "icp4r_register_map_batch", "icp4r_register_batch", "icp4r_shard_unique_id", "icp4r_shard_init", "icp4r_shard_ipc_export", "icp4r_shard_ipc_import",
    "icp4r_register_sharded", "icp4r_accumulate_slab", "icp4r_register_submap", "icp4r_doppler_static_points", "icp4r_transform_points", "icp4r_voxel_grid", "icp4r_odometry_step", "icp4r_map_box_search",
    "icp4r_map_radius_search", "icp4r_map_delete_boxes", "icp4r_map_add_boxes", "icp4r_map_delete_points", "icp4r_doppler_filter",
]


class Opts(C.Structure):
    _fields_ = [
        ("residual", C.c_int32),
        ("k", C.c_int32),
        ("max_iterations", C.c_int32),
        ("early_exit", C.c_int32),
        ("max_corr_dist", C.c_double),
        ("rot_eps", C.c_double),
        ("trans_eps", C.c_double),
        ("mse_abs_eps", C.c_double),
        ("plane_thresh", C.c_double),
        ("T0", C.c_double * 16),
        ("interp_s", C.c_double),
    ]


class Result(C.Structure):
    _fields_ = [
        ("converged", C.c_int32),
        ("iterations", C.c_int32),
        ("n_corr", C.c_int32),
        ("n_fitness", C.c_int32),
        ("fitness", C.c_double),
        ("last_cost", C.c_double),
    ]


class DopplerOpts(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("reserved", C.c_int32), ("seed", C.c_uint64), ("sigma", C.c_double), ("split", C.c_double)]


class DopplerResult(C.Structure):
    _fields_ = [("A", C.c_double), ("b", C.c_double), ("score", C.c_double), ("velocity", C.c_double * 3), ("n_static", C.c_int32),
                ("best_iteration", C.c_int32)]


class Dump(C.Structure):
    _fields_ = [("pose", C.c_void_p), ("acc", C.c_void_p), ("idx", C.c_void_p)]


RESULT_DTYPE = np.dtype([("converged", np.int32), ("iterations", np.int32), ("n_corr", np.int32), ("n_fitness", np.int32),
                         ("fitness", np.float64), ("last_cost", np.float64)])
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


class Icp4rError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"icp4r status {code}: {msg}")
        self.code = code


def lib_path() -> str:
    return os.path.join(_HERE, "libicp4r_cuda.so")


_lib = None


def load_library() -> C.CDLL:
    """Load libicp4r_cuda.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise FileNotFoundError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(there is no CPU fallback)")
        lib = C.CDLL(p)
        lib.icp4r_last_error.restype = C.c_char_p
        lib.icp4r_last_error.argtypes = [C.c_void_p]
        lib.icp4r_version.restype = C.c_char_p
        _lib = lib
    return _lib


def default_opts(**kw) -> Opts:
    o = Opts()
    load_library().icp4r_default_opts(C.byref(o))
    for k, v in kw.items():
        if k == "T0":
            v = np.asarray(v, dtype=np.float64).reshape(16)
            for i in range(16):
                o.T0[i] = float(v[i])
        else:
            setattr(o, k, v)
    return o


def _ptr(a):
    """(pointer, mem) for a numpy array, a torch tensor (cpu or cuda) or None."""
    if a is None:
        return None, None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data), HOST
    # torch tensor
    return C.c_void_p(a.data_ptr()), (DEVICE if a.is_cuda else HOST)


def _f4(a):
    """point rows as float32: (n, 4) packed x y z w by default; wider (or 3-float) rows go with Icp4r.set_point_layout"""
    if isinstance(a, np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.float32)
        assert a.ndim == 2 and a.shape[1] >= 3, a.shape
        return a
    assert a.dim() == 2 and a.shape[1] >= 3 and a.is_contiguous() and a.dtype.is_floating_point and a.element_size() == 4
    return a


def _knn_k(residual, k):
    return 1 if residual in (P2P_SVD, P2P_GN, GICP) else (2 if residual == P2LINE else (3 if residual == P2PLANE_3PT else (k if k > 0 else 5)))


class Icp4r:
    """One handle = one device + one stream. Methods take numpy arrays (host path, copies inside the call) or
    torch CUDA tensors (device path, zero-copy)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.icp4r_create(C.c_int(device), C.byref(h))
        if rc != 0:
            raise Icp4rError(rc, self.lib.icp4r_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.icp4r_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise Icp4rError(rc, self.lib.icp4r_last_error(self.h).decode())

    # ---- lifecycle
    def set_stream(self, cuda_stream_ptr: int):
        self._ck(self.lib.icp4r_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_point_layout(self, stride_bytes: int = 16, w_offset_bytes: int = 12):
        """layout of the point INPUT rows of later calls (pcl::PointXYZI: 32, 16); outputs stay packed (n, 4)"""
        self._ck(self.lib.icp4r_set_point_layout(self.h, C.c_int32(stride_bytes), C.c_int32(w_offset_bytes)))

    def synchronize(self):
        self._ck(self.lib.icp4r_synchronize(self.h))

    def launch_count(self) -> int:
        v = C.c_int64(0)
        self._ck(self.lib.icp4r_launch_count(self.h, C.byref(v)))
        return v.value

    def set_profiling(self, on: bool):
        self._ck(self.lib.icp4r_set_profiling(self.h, C.c_int(int(on))))

    def last_profile(self):
        buf = np.zeros(4096, np.float32)
        n = C.c_int32(0)
        self._ck(self.lib.icp4r_last_profile(self.h, C.c_void_p(buf.ctypes.data), C.c_int32(buf.shape[0]), C.byref(n)))
        return buf[:n.value].copy()

    def set_stats(self, on: bool):
        self._ck(self.lib.icp4r_set_stats(self.h, C.c_int(int(on))))

    def get_stats(self):
        """[searches, squared-distance evaluations, points settled without a search, ...] since the last call"""
        out = (C.c_int64 * 8)()
        self._ck(self.lib.icp4r_get_stats(self.h, out))
        return [int(v) for v in out]

    # ---- map
    def map_build(self, pts, cell_size: float = 0.0):
        pts = _f4(pts)
        p, mem = _ptr(pts)
        self._ck(self.lib.icp4r_map_build(self.h, p, C.c_int32(pts.shape[0]), C.c_int(mem), C.c_float(cell_size)))

    def map_set_downsample(self, voxel: float):
        self._ck(self.lib.icp4r_map_set_downsample(self.h, C.c_float(voxel)))

    def map_add_points(self, pts, downsample_on: bool = False) -> int:
        pts = _f4(pts)
        p, mem = _ptr(pts)
        r = C.c_int32(0)
        self._ck(self.lib.icp4r_map_add_points(self.h, p, C.c_int32(pts.shape[0]), C.c_int(mem), C.c_int(int(downsample_on)),
                                               C.byref(r)))
        return r.value

    def map_size(self):
        s, v = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.icp4r_map_size(self.h, C.byref(s), C.byref(v)))
        return s.value, v.value

    def map_range(self):
        out = np.zeros(6, np.float32)
        self._ck(self.lib.icp4r_map_range(self.h, C.c_void_p(out.ctypes.data)))
        return out

    def _knn(self, fn, q, k, max_dist, out=None):
        q = _f4(q)
        p, mem = _ptr(q)
        nq = q.shape[0]
        if mem == HOST:
            idx = np.empty((nq, k), np.int32)
            d2 = np.empty((nq, k), np.float32)
            found = np.empty(nq, np.int32)
        else:
            import torch
            if out is not None:
                idx, d2, found = out
            else:
                idx = torch.empty((nq, k), dtype=torch.int32, device=q.device)
                d2 = torch.empty((nq, k), dtype=torch.float32, device=q.device)
                found = torch.empty(nq, dtype=torch.int32, device=q.device)
        self._ck(fn(self.h, p, C.c_int32(nq), C.c_int(mem), C.c_int32(k), C.c_double(max_dist), _ptr(idx)[0], _ptr(d2)[0],
                    _ptr(found)[0]))
        return idx, d2, found

    def map_knn(self, q, k, max_dist=0.0, out=None):
        return self._knn(self.lib.icp4r_map_knn, q, k, max_dist, out)

    def map_knn_brute(self, q, k, max_dist=0.0, out=None):
        return self._knn(self.lib.icp4r_map_knn_brute, q, k, max_dist, out)

    def map_points(self):
        n, _ = self.map_size()
        pts = np.zeros((n, 4), np.float32)
        valid = np.zeros(n, np.uint8)
        self._ck(self.lib.icp4r_map_points(self.h, C.c_int(HOST), C.c_void_p(pts.ctypes.data), C.c_void_p(valid.ctypes.data),
                                           C.c_int32(n)))
        return pts, valid

    def map_points_dev(self):
        """the stored points in insertion order as a torch tensor on the handle's device"""
        import torch
        n, _ = self.map_size()
        pts = torch.zeros((max(n, 1), 4), dtype=torch.float32, device=torch.device("cuda", self.device))
        self._ck(self.lib.icp4r_map_points(self.h, C.c_int(DEVICE), C.c_void_p(pts.data_ptr()), None, C.c_int32(n)))
        self.synchronize()
        return pts[:n]

    def map_sector(self, centre, radius, heading_deg):
        n, _ = self.map_size()
        c = np.asarray(centre, np.float32)
        out = np.empty(max(n, 1), np.int32)
        cnt = C.c_int32(0)
        self._ck(self.lib.icp4r_map_sector(self.h, C.c_void_p(c.ctypes.data), C.c_float(radius), C.c_float(heading_deg),
                                           C.c_int(HOST), C.c_void_p(out.ctypes.data), C.c_int32(out.shape[0]), C.byref(cnt)))
        return out[:min(cnt.value, out.shape[0])].copy()

    def map_box_search(self, bmin, bmax):
        n, _ = self.map_size()
        lo, hi = np.ascontiguousarray(bmin, np.float32), np.ascontiguousarray(bmax, np.float32)
        out = np.empty(max(n, 1), np.int32)
        cnt = C.c_int32(0)
        self._ck(self.lib.icp4r_map_box_search(self.h, C.c_void_p(lo.ctypes.data), C.c_void_p(hi.ctypes.data), C.c_int(HOST),
                                               C.c_void_p(out.ctypes.data), C.c_int32(out.shape[0]), C.byref(cnt)))
        return out[:min(cnt.value, out.shape[0])].copy()

    def map_radius_search(self, centre, radius):
        n, _ = self.map_size()
        c = np.ascontiguousarray(centre, np.float32)
        out = np.empty(max(n, 1), np.int32)
        cnt = C.c_int32(0)
        self._ck(self.lib.icp4r_map_radius_search(self.h, C.c_void_p(c.ctypes.data), C.c_float(radius), C.c_int(HOST),
                                                  C.c_void_p(out.ctypes.data), C.c_int32(out.shape[0]), C.byref(cnt)))
        return out[:min(cnt.value, out.shape[0])].copy()

    def map_delete_boxes(self, boxes) -> int:
        b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        cnt = C.c_int32(0)
        self._ck(self.lib.icp4r_map_delete_boxes(self.h, C.c_void_p(b.ctypes.data), C.c_int32(b.shape[0]), C.byref(cnt)))
        return cnt.value

    def map_add_boxes(self, boxes) -> int:
        b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        cnt = C.c_int32(0)
        self._ck(self.lib.icp4r_map_add_boxes(self.h, C.c_void_p(b.ctypes.data), C.c_int32(b.shape[0]), C.byref(cnt)))
        return cnt.value

    def map_delete_points(self, pts) -> int:
        pts = _f4(pts)
        p, mem = _ptr(pts)
        cnt = C.c_int32(0)
        self._ck(self.lib.icp4r_map_delete_points(self.h, p, C.c_int32(pts.shape[0]), C.c_int(mem), C.byref(cnt)))
        return cnt.value

    # ---- registration
    def _dump_bufs(self, opts, n, dump, mem, device=None):
        if not dump:
            return None, None
        it = max(opts.max_iterations, 0)
        k = _knn_k(opts.residual, opts.k)
        if mem == HOST:
            dp = np.zeros((it, 16), np.float64)
            da = np.zeros((it, ACC_LEN), np.float64)
            di = np.full((it, n, k), -1, np.int32)
        else:
            import torch
            dp = torch.zeros((it, 16), dtype=torch.float64, device=device)
            da = torch.zeros((it, ACC_LEN), dtype=torch.float64, device=device)
            di = torch.full((it, n, k), -1, dtype=torch.int32, device=device)
        d = Dump(_ptr(dp)[0], _ptr(da)[0], _ptr(di)[0])
        return d, (dp, da, di)

    def register(self, src, tgt, opts: Opts, dump: bool = False):
        src, tgt = _f4(src), _f4(tgt)
        ps, mem = _ptr(src)
        pt, mem2 = _ptr(tgt)
        assert mem == mem2
        T = np.zeros(16, np.float64)
        res = Result()
        d, bufs = self._dump_bufs(opts, src.shape[0], dump, mem, getattr(src, "device", None))
        self._ck(self.lib.icp4r_register(self.h, ps, C.c_int32(src.shape[0]), pt, C.c_int32(tgt.shape[0]), C.c_int(mem),
                                         C.byref(opts), C.c_void_p(T.ctypes.data), C.byref(res),
                                         C.byref(d) if d is not None else None))
        return T.reshape(4, 4), res, bufs

    def register_map(self, src, opts: Opts, dump: bool = False):
        src = _f4(src)
        ps, mem = _ptr(src)
        T = np.zeros(16, np.float64)
        res = Result()
        d, bufs = self._dump_bufs(opts, src.shape[0], dump, mem, getattr(src, "device", None))
        self._ck(self.lib.icp4r_register_map(self.h, ps, C.c_int32(src.shape[0]), C.c_int(mem), C.byref(opts),
                                             C.c_void_p(T.ctypes.data), C.byref(res), C.byref(d) if d is not None else None))
        return T.reshape(4, 4), res, bufs

    def register_map_batch(self, src, off, opts: Opts, T0s=None):
        """several scans (concatenated [sum n, 4], numpy or CUDA tensor; off int32 [B+1]) against the map at once.
        Returns (T [B,4,4], results structured array)."""
        src = _f4(src)
        ps, mem = _ptr(src)
        off = np.ascontiguousarray(off, np.int32)
        B = off.shape[0] - 1
        T = np.zeros((B, 16), np.float64)
        res = np.zeros(B, RESULT_DTYPE)
        t0 = None if T0s is None else np.ascontiguousarray(T0s, np.float64).reshape(B, 16)
        self._ck(self.lib.icp4r_register_map_batch(self.h, ps, C.c_void_p(off.ctypes.data), C.c_int32(B), C.c_int(mem), C.byref(opts),
                                                   None if t0 is None else C.c_void_p(t0.ctypes.data), C.c_void_p(T.ctypes.data),
                                                   C.c_void_p(res.ctypes.data)))
        return T.reshape(B, 4, 4), res

    def register_batch(self, src, src_off, tgt, tgt_off, opts: Opts, out=None):
        """src/tgt: concatenated clouds [sum n, 4]; *_off: int32 [n_pairs+1]. Returns (T [P,4,4], results)."""
        src, tgt = _f4(src), _f4(tgt)
        ps, mem = _ptr(src)
        pt, _ = _ptr(tgt)
        n_pairs = int(src_off.shape[0]) - 1
        if mem == HOST:
            src_off = np.ascontiguousarray(src_off, np.int32)
            tgt_off = np.ascontiguousarray(tgt_off, np.int32)
            T = np.zeros((n_pairs, 16), np.float64)
            res = np.zeros(n_pairs, RESULT_DTYPE)
        else:
            import torch
            if out is not None:
                T, res = out
            else:
                T = torch.zeros((n_pairs, 16), dtype=torch.float64, device=src.device)
                res = torch.zeros((n_pairs, RESULT_DTYPE.itemsize), dtype=torch.uint8, device=src.device)
        self._ck(self.lib.icp4r_register_batch(self.h, ps, _ptr(src_off)[0], pt, _ptr(tgt_off)[0], C.c_int32(n_pairs), C.c_int(mem),
                                               C.byref(opts), _ptr(T)[0], _ptr(res)[0]))
        if mem == HOST:
            return T.reshape(n_pairs, 4, 4), res
        return T, res

    # ---- sharding
    @staticmethod
    def shard_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = load_library().icp4r_shard_unique_id(buf)
        if rc != 0:
            raise Icp4rError(rc, "icp4r_shard_unique_id failed (NCCL not loadable?)")
        return buf.raw

    def shard_init(self, uid: bytes, rank: int, world: int):
        assert len(uid) == 128
        self._ck(self.lib.icp4r_shard_init(self.h, C.c_char_p(uid), C.c_int(rank), C.c_int(world)))

    def shard_ipc_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.lib.icp4r_shard_ipc_export(self.h, buf))
        return buf.raw

    def shard_ipc_import(self, handles, rank: int, world: int):
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        self._ck(self.lib.icp4r_shard_ipc_import(self.h, C.c_char_p(blob), C.c_int(rank), C.c_int(world)))

    def register_sharded(self, src, opts: Opts, axis: int, slab_lo: float, slab_hi: float):
        src = _f4(src)
        ps, mem = _ptr(src)
        T = np.zeros(16, np.float64)
        res = Result()
        self._ck(self.lib.icp4r_register_sharded(self.h, ps, C.c_int32(src.shape[0]), C.c_int(mem), C.byref(opts), C.c_int(axis),
                                                 C.c_float(slab_lo), C.c_float(slab_hi), C.c_void_p(T.ctypes.data), C.byref(res)))
        return T.reshape(4, 4), res

    # ---- Doppler filter
    def register_submap(self, src, idx, opts: Opts):
        """source vs the map points idx (numpy int32 or a torch int32 tensor on the device, like src). Returns (T, result)."""
        src = _f4(src)
        ps, mem = _ptr(src)
        pi, mem_i = _ptr(idx)
        assert mem == mem_i, "src and idx must live in the same memory space"
        T = np.zeros(16, np.float64)
        res = Result()
        self._ck(self.lib.icp4r_register_submap(self.h, ps, C.c_int32(src.shape[0]), pi, C.c_int32(int(idx.shape[0])), C.c_int(mem), C.byref(opts),
                                                C.c_void_p(T.ctypes.data), C.byref(res)))
        return T.reshape(4, 4), res

    def map_sector_dev(self, centre, radius, heading_deg):
        """Sector_Search with the indices left on the device (torch int32 tensor)"""
        import torch
        n, _ = self.map_size()
        c = np.asarray(centre, np.float32)
        out = torch.empty(max(n, 1), dtype=torch.int32, device=torch.device("cuda", self.device))
        cnt = C.c_int32(0)
        self._ck(self.lib.icp4r_map_sector(self.h, C.c_void_p(c.ctypes.data), C.c_float(radius), C.c_float(heading_deg),
                                           C.c_int(DEVICE), C.c_void_p(out.data_ptr()), C.c_int32(out.shape[0]), C.byref(cnt)))
        return out[:min(cnt.value, out.shape[0])]

    def doppler_static_points(self, records, iterations: int = 0, seed: int = 1, sigma: float = 0.5, split: float = 0.2):
        """(static points [ns,4] in the memory space of `records`, DopplerResult)"""
        o = DopplerOpts(iterations, 0, seed, sigma, split)
        res = DopplerResult()
        pr, mem = _ptr(records)
        n = int(records.shape[0])
        cnt = C.c_int32(0)
        if mem == HOST:
            rec = np.ascontiguousarray(records, np.float32)
            out = np.empty((max(n, 1), 4), np.float32)
            self._ck(self.lib.icp4r_doppler_static_points(self.h, C.c_void_p(rec.ctypes.data), C.c_int32(n), C.c_int(HOST), C.byref(o),
                                                          C.c_void_p(out.ctypes.data), C.c_int32(n), C.byref(cnt), C.byref(res)))
            return out[:cnt.value].copy(), res
        import torch
        out = torch.empty((max(n, 1), 4), dtype=torch.float32, device=records.device)
        self._ck(self.lib.icp4r_doppler_static_points(self.h, pr, C.c_int32(n), C.c_int(DEVICE), C.byref(o), C.c_void_p(out.data_ptr()),
                                                      C.c_int32(n), C.byref(cnt), C.byref(res)))
        return out[:cnt.value], res

    def accumulate_slab(self, src, opts: Opts, T, axis: int = -1, slab_lo: float = 0.0, slab_hi: float = 0.0):
        """this slab's partial accumulators [ACC_LEN] at pose T (no cross-rank sum, no solve)"""
        src = _f4(src)
        ps, mem = _ptr(src)
        Tm = np.ascontiguousarray(T, np.float64).reshape(16)
        acc = np.zeros(ACC_LEN, np.float64)
        self._ck(self.lib.icp4r_accumulate_slab(self.h, ps, C.c_int32(src.shape[0]), C.c_int(mem), C.byref(opts), C.c_void_p(Tm.ctypes.data),
                                                C.c_int(axis), C.c_float(slab_lo), C.c_float(slab_hi), C.c_void_p(acc.ctypes.data)))
        return acc

    def doppler_filter(self, records, iterations: int = 0, seed: int = 1, sigma: float = 0.5, split: float = 0.2):
        """records: [n,5] x,y,z,intensity,v_r (numpy or CUDA tensor). Returns (static_mask uint8 [n], DopplerResult)."""
        if isinstance(records, np.ndarray):
            records = np.ascontiguousarray(records, np.float32)
            mask = np.zeros(records.shape[0], np.uint8)
        else:
            import torch
            mask = torch.zeros(records.shape[0], dtype=torch.uint8, device=records.device)
        p, mem = _ptr(records)
        o = DopplerOpts(iterations, 0, seed, sigma, split)
        r = DopplerResult()
        self._ck(self.lib.icp4r_doppler_filter(self.h, p, C.c_int32(records.shape[0]), C.c_int(mem), C.byref(o), _ptr(mask)[0], C.byref(r)))
        return mask, r

    # ---- helpers
    def odometry_step(self, scan, opts: Opts, T_prior):
        """register against the map from T_prior, then insert the scan at the estimated pose; returns (T, result)"""
        scan = _f4(scan)
        p, mem = _ptr(scan)
        T = np.ascontiguousarray(T_prior, np.float64).reshape(16).copy()
        res = Result()
        self._ck(self.lib.icp4r_odometry_step(self.h, p, C.c_int32(scan.shape[0]), C.c_int(mem), C.byref(opts), C.c_int(0),
                                              C.c_void_p(T.ctypes.data), C.byref(res)))
        return T.reshape(4, 4), res

    def voxel_grid(self, pts, leaf: float, out=None):
        """pcl::VoxelGrid centroid filter. pts None: the handle's map (returns numpy, or — with `out`, a torch CUDA float32
        [cap,4] tensor — the filled part of `out`); numpy -> numpy; torch CUDA -> torch."""
        cnt = C.c_int32(0)
        if pts is None and out is not None:
            self._ck(self.lib.icp4r_voxel_grid(self.h, None, C.c_int32(0), C.c_int(DEVICE), C.c_float(leaf), C.c_void_p(out.data_ptr()),
                                               C.c_int32(out.shape[0]), C.byref(cnt)))
            return out[:min(cnt.value, out.shape[0])]
        if pts is None:
            n, _ = self.map_size()
            out = np.empty((max(n, 1), 4), np.float32)
            self._ck(self.lib.icp4r_voxel_grid(self.h, None, C.c_int32(0), C.c_int(HOST), C.c_float(leaf), C.c_void_p(out.ctypes.data),
                                               C.c_int32(out.shape[0]), C.byref(cnt)))
            return out[:cnt.value].copy()
        pts = _f4(pts)
        p, mem = _ptr(pts)
        if mem == HOST:
            out = np.empty((max(pts.shape[0], 1), 4), np.float32)
        else:
            import torch
            out = torch.empty((max(pts.shape[0], 1), 4), dtype=torch.float32, device=pts.device)
        self._ck(self.lib.icp4r_voxel_grid(self.h, p, C.c_int32(pts.shape[0]), C.c_int(mem), C.c_float(leaf), _ptr(out)[0],
                                           C.c_int32(out.shape[0]), C.byref(cnt)))
        return out[:cnt.value].copy() if mem == HOST else out[:cnt.value]

    def voxel_grid_into(self, pts, leaf: float, out) -> int:
        """voxel_grid of a CUDA tensor into a caller-provided CUDA float32 [cap, 4] tensor; returns the number of leaves"""
        pts = _f4(pts)
        cnt = C.c_int32(0)
        self._ck(self.lib.icp4r_voxel_grid(self.h, C.c_void_p(pts.data_ptr()), C.c_int32(pts.shape[0]), C.c_int(DEVICE), C.c_float(leaf),
                                           C.c_void_p(out.data_ptr()), C.c_int32(out.shape[0]), C.byref(cnt)))
        return cnt.value

    def transform_points(self, T, pts):
        pts = _f4(pts)
        p, mem = _ptr(pts)
        T = np.ascontiguousarray(T, np.float64).reshape(16)
        if mem == HOST:
            out = np.empty((pts.shape[0], 4), np.float32)  # outputs are always packed x, y, z, w
        else:
            import torch
            out = torch.empty((pts.shape[0], 4), dtype=torch.float32, device=pts.device)
        self._ck(self.lib.icp4r_transform_points(self.h, C.c_void_p(T.ctypes.data), p, C.c_int32(pts.shape[0]), C.c_int(mem),
                                                 _ptr(out)[0]))
        return out
