"""GPU parity: Doppler static-point filter + ego velocity (fitSineRansac / split / least squares of
/root/reference/src/iterative_closest_point.cpp:85-128,387-431) against the oracle with the same seeded hypotheses."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_frame(pkg, seed, n, v_ego, dynamic_frac=0.1):
    s = pkg.synth
    rng = np.random.default_rng(seed)
    pts = s.Scene(seed).sample(rng, n, radius=60.0)
    vr, dyn = s.doppler(rng, pts, v_ego, dynamic_frac)
    return s.radar_frame_bin(pts, vr), dyn


@pytest.mark.parametrize("seed,n,iters", [(3, 4000, 0), (4, 1500, 64), (5, 333, 500)])
def test_doppler_filter_matches_oracle(pkg, O, handle, seed, n, iters):
    rec, dyn = make_frame(pkg, seed, n, np.array([3.0, 0.4, 0.0]))
    mask, r = handle.doppler_filter(rec, iters, seed=seed + 100)
    omask, o = O.doppler_filter(rec, iters, seed=seed + 100)
    assert r.best_iteration == o.best_iteration and r.score == o.score      # same winning hypothesis, same inlier count
    assert (mask == omask).all() and r.n_static == o.n_static == int(omask.sum())
    assert abs(r.A - o.A) <= 1e-12 * abs(o.A) and abs(r.b - o.b) <= 1e-12
    assert np.allclose(np.array(list(r.velocity)), np.array(list(o.v)), rtol=1e-9, atol=1e-12)
    # sanity of the synthetic frame: the sine amplitude is the ego speed in the radar plane
    assert abs(abs(r.A) - np.hypot(3.0, 0.4)) < 0.15


def test_doppler_edge_cases(pkg, O, handle):
    rec, _ = make_frame(pkg, 9, 64, np.array([0.0, 0.0, 0.0]), dynamic_frac=0.0)   # stationary sensor
    mask, r = handle.doppler_filter(rec, 0, seed=1)
    omask, o = O.doppler_filter(rec, 0, seed=1)
    assert (mask == omask).all() and r.best_iteration == o.best_iteration
    # empty frame
    mask, r = handle.doppler_filter(np.zeros((0, 5), np.float32), 0, seed=1)
    assert mask.shape == (0,) and r.n_static == 0 and r.best_iteration == -1
    # device-resident records
    import torch
    rec, _ = make_frame(pkg, 10, 2000, np.array([1.0, -2.0, 0.0]))
    m1, r1 = handle.doppler_filter(rec, 0, seed=5)
    m2, r2 = handle.doppler_filter(torch.from_numpy(rec).cuda(), 0, seed=5)
    assert (m2.cpu().numpy() == m1).all() and r1.best_iteration == r2.best_iteration
