"""Host-side logic for slab-sharded maps: split a map into spatial slabs along one
axis with a halo, so that each GPU holds one slab and the exact gated neighbours of every source point it owns.

Ownership rule (device side, reg_iter_kernel): a source point belongs to the rank whose half-open interval
[slab_lo, slab_hi) contains its *current transformed* coordinate along the axis; the outermost slabs extend to
-inf / +inf.  Halo guarantee: rank r stores every map point with coordinate in [slab_lo - halo, slab_hi + halo],
so for halo >= max_corr_dist (plus float slack) all points within the gate of an owned query are local and the
union over ranks of the per-rank accumulators equals the single-map accumulators.
"""
from __future__ import annotations

import numpy as np


def slab_bounds(coord: np.ndarray, world: int) -> np.ndarray:
    """world+1 boundaries (first -inf, last +inf) giving each slab about the same number of map points."""
    qs = np.quantile(coord.astype(np.float64), np.linspace(0, 1, world + 1)[1:-1]) if world > 1 else np.array([])
    b = np.concatenate([[-np.inf], qs, [np.inf]]).astype(np.float32)
    return b


def slab_of_rank(pts: np.ndarray, rank: int, world: int, axis: int = 0, halo: float = 2.0, bounds: np.ndarray | None = None):
    """(points of this rank's slab incl. halo, slab_lo, slab_hi, global indices of those points)."""
    c = pts[:, axis]
    b = slab_bounds(c, world) if bounds is None else bounds
    lo, hi = float(b[rank]), float(b[rank + 1])
    h = np.float32(halo) * np.float32(1.0 + 1e-5) + np.float32(1e-4)
    keep = (c >= np.float32(lo) - h) & (c <= np.float32(hi) + h)
    idx = np.nonzero(keep)[0]
    return np.ascontiguousarray(pts[idx]), lo, hi, idx


def owner_of(coord: np.ndarray, bounds: np.ndarray) -> np.ndarray:
    """rank owning each (float32) coordinate under the half-open rule used on the device"""
    return np.clip(np.searchsorted(bounds[1:-1], coord.astype(np.float32), side="right"), 0, len(bounds) - 2)


def pair_range(n_pairs: int, rank: int, world: int):
    """contiguous pair range [lo, hi) of a rank for batched registration (C4): independent units, no collective"""
    lo = (n_pairs * rank) // world
    hi = (n_pairs * (rank + 1)) // world
    return lo, hi
