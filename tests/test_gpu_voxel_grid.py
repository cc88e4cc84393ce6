"""GPU parity: centroid-per-leaf voxel filter (icp4r_voxel_grid) against the oracle's restatement of pcl::VoxelGrid —
bit-exact centroids in the same (ascending leaf) order — plus size-independent properties at map scale."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


def cloud(seed, n, extent=40.0, flat=0.2):
    rng = np.random.default_rng(seed)
    p = np.zeros((n, 4), np.float32)
    p[:, :3] = rng.uniform(-extent, extent, (n, 3)) * np.array([1, 1, flat])
    p[:, 3] = rng.uniform(0, 50, n)
    return p


@pytest.mark.parametrize("leaf", [0.5, 0.3, 1.7])
@pytest.mark.parametrize("n", [1, 33, 5000, 70000])
def test_voxel_grid_matches_oracle(handle, O, leaf, n):
    p = cloud(n * 7 + int(leaf * 10), n)
    got = handle.voxel_grid(p, leaf)
    want = O.voxel_grid(p, leaf)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert (bits(got) == bits(want)).all()


def test_voxel_grid_edge_cases(handle, O):
    rng = np.random.default_rng(5)
    p = cloud(11, 4000, extent=6.0)
    p[::97, 0] = np.nan                      # non-finite points are skipped
    p[5::131, 2] = np.inf
    p[1000:1200, :3] = (rng.integers(-12, 12, (200, 3)) * np.float32(0.5)).astype(np.float32)  # exactly on leaf faces
    p[2000:2300] = p[2000]                   # 300 copies in one leaf
    p[3000:3050, :3] = -0.0                 # signed zeros
    got = handle.voxel_grid(p, 0.5)
    want = O.voxel_grid(p, 0.5)
    assert got.shape == want.shape and (bits(got) == bits(want)).all()
    # empty input, all-NaN input
    assert handle.voxel_grid(np.zeros((0, 4), np.float32), 0.5).shape == (0, 4)
    assert handle.voxel_grid(np.full((10, 4), np.nan, np.float32), 0.5).shape == (0, 4)
    # torch CUDA tensors: zero-copy path gives the same bits
    import torch
    d = handle.voxel_grid(torch.from_numpy(p).cuda(), 0.5)
    assert (bits(d.cpu().numpy()) == bits(want)).all()
    # a leaf so small that the index space overflows is refused, with a message
    from icp4r_loader import pkg
    with pytest.raises(pkg.Icp4rError):
        handle.voxel_grid(cloud(1, 100, extent=4000.0, flat=1.0), 0.001)


def test_voxel_grid_of_the_map_skips_deleted_points(handle, O):
    base = cloud(21, 6000, extent=8.0)
    handle.map_build(base)
    handle.map_set_downsample(0.5)
    handle.map_add_points(cloud(22, 3000, extent=8.0), True)   # voxel down-sampling deletes points
    pts, valid = handle.map_points()
    assert (valid == 0).any()
    got = handle.voxel_grid(None, 0.8)
    want = O.voxel_grid(pts, 0.8, valid=valid)
    assert got.shape == want.shape and (bits(got) == bits(want)).all()


def test_voxel_grid_large_properties(handle):
    """6 M points (the C3 end-of-sequence map size): leaves == distinct leaf indices, order ascending, count-weighted
    centroids reproduce the cloud's mean, idempotent on its own output when every output stays in its leaf"""
    import torch
    n = 6_000_000
    g = torch.Generator(device="cuda").manual_seed(3)
    p = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    p[:, 0] = (torch.rand(n, generator=g, device="cuda") - 0.5) * 400
    p[:, 1] = (torch.rand(n, generator=g, device="cuda") - 0.5) * 400
    p[:, 2] = (torch.rand(n, generator=g, device="cuda") - 0.5) * 20
    p[:, 3] = 1.0
    leaf = 0.5
    out = handle.voxel_grid(p, leaf)
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = torch.floor(p[:, :3] * float(inv)).to(torch.int64)
    mn = ijk.min(0).values
    div = ijk.max(0).values - mn + 1
    key = (ijk[:, 0] - mn[0]) + (ijk[:, 1] - mn[1]) * div[0] + (ijk[:, 2] - mn[2]) * div[0] * div[1]
    uk, cnt = torch.unique(key, return_counts=True)
    assert out.shape[0] == uk.numel()
    # every centroid lies in its leaf (up to float rounding of the mean) and leaves come in ascending order
    oj = torch.floor(out[:, :3] * float(inv)).to(torch.int64)
    okey = (oj[:, 0] - mn[0]) + (oj[:, 1] - mn[1]) * div[0] + (oj[:, 2] - mn[2]) * div[0] * div[1]
    assert (okey == uk).float().mean().item() > 0.9999
    mean_from_leaves = (out[:, :3].double() * cnt[:, None].double()).sum(0) / n
    assert torch.allclose(mean_from_leaves, p[:, :3].double().mean(0), atol=1e-4)
    assert torch.allclose(out[:, 3], torch.ones_like(out[:, 3]))
