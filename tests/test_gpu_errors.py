"""GPU: error behaviour at the C boundary — status codes, messages, no exceptions, state checks."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_status_codes_and_messages(pkg):
    h = pkg.Icp4r(0)
    lib = h.lib
    q = np.zeros((4, 4), np.float32)
    # map queries before Build -> ICP4R_ERR_STATE with a message
    with pytest.raises(pkg.Icp4rError) as e:
        h.map_knn(q, 1)
    assert e.value.code == 4 and "before" in str(e.value)
    with pytest.raises(pkg.Icp4rError) as e:
        h.register_map(q, pkg.default_opts())
    assert e.value.code == 4
    h.map_build(np.random.default_rng(0).normal(size=(100, 4)).astype(np.float32))
    # bad k, bad mem flag, null pointers -> ICP4R_ERR_INVALID
    with pytest.raises(pkg.Icp4rError) as e:
        h.map_knn(q, 0)
    assert e.value.code == 1
    with pytest.raises(pkg.Icp4rError) as e:
        h.map_knn(q, 17)
    assert e.value.code == 1
    idx = np.zeros((4, 1), np.int32)
    d2 = np.zeros((4, 1), np.float32)
    assert lib.icp4r_map_knn(h.h, C.c_void_p(q.ctypes.data), C.c_int32(4), C.c_int(7), C.c_int32(1), C.c_double(0.0),
                             C.c_void_p(idx.ctypes.data), C.c_void_p(d2.ctypes.data), None) == 1
    assert lib.icp4r_map_knn(h.h, None, C.c_int32(4), C.c_int(0), C.c_int32(1), C.c_double(0.0), C.c_void_p(idx.ctypes.data),
                             C.c_void_p(d2.ctypes.data), None) == 1
    assert b"bad arguments" in lib.icp4r_last_error(h.h)
    # unsupported / invalid registration options
    with pytest.raises(pkg.Icp4rError) as e:
        h.register_map(q, pkg.default_opts(residual=9))
    assert e.value.code == 1
    with pytest.raises(pkg.Icp4rError) as e:
        h.register_map(q, pkg.default_opts(residual=pkg.P2PLANE_KNN, k=2))
    assert e.value.code == 1
    with pytest.raises(pkg.Icp4rError) as e:
        h.register_batch(q, np.array([0, 4], np.int32), q, np.array([0, 4], np.int32), pkg.default_opts(residual=pkg.P2PLANE_KNN))
    assert e.value.code == 6
    with pytest.raises(pkg.Icp4rError) as e:
        h.register_sharded(q, pkg.default_opts(residual=pkg.P2PLANE_KNN, max_corr_dist=0.0), 0, -1.0, 1.0)
    assert e.value.code == 1 and "max_corr_dist" in str(e.value)
    # a failed call leaves the handle usable
    i2, _, _ = h.map_knn(q, 1)
    assert (i2 >= 0).all()
    # device index out of range at creation
    hh = C.c_void_p()
    assert lib.icp4r_create(C.c_int(1 << 20), C.byref(hh)) == 1 and not hh.value
    h.close()


def test_launch_counter_and_determinism(pkg):
    """the same call twice gives bit-identical poses (fixed summation order) and the launch counter advances"""
    h = pkg.Icp4r(0)
    src, tgt, _ = pkg.synth.frame_pair(3, 2000, 9000, extent=30.0)
    h.map_build(tgt)
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=7, max_corr_dist=2.0)
    n0 = h.launch_count()
    T1, r1, _ = h.register_map(src, o)
    n1 = h.launch_count()
    T2, r2, _ = h.register_map(src, o)
    assert np.array_equal(T1, T2) and r1.fitness == r2.fitness
    assert n1 - n0 == 1 + 7 + 1 and h.launch_count() - n1 == 9    # state init + 7 iterations + fitness pass
    h.close()


def test_profiling_mode_with_gicp_does_not_crash(handle, pkg):
    """per-launch profiling is only produced for the plain iteration loop; GICP must simply return an empty profile"""
    import numpy as np
    scan, mp, _ = pkg.synth.scan_to_map(5, 800, 20000, extent=40.0, scan_radius=30.0)
    handle.map_build(mp)
    handle.set_profiling(True)
    try:
        o = pkg.default_opts(residual=pkg.GICP, k=5, max_iterations=5, max_corr_dist=2.0)
        T, res, _ = handle.register_map(scan, o)
        assert np.isfinite(T).all() and len(handle.last_profile()) == 0
        o = pkg.default_opts(residual=pkg.P2P_GN, max_iterations=5, max_corr_dist=2.0)
        handle.register_map(scan, o)
        assert len(handle.last_profile()) == 6
    finally:
        handle.set_profiling(False)
