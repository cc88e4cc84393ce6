"""Seeded synthetic 4D-radar scenes for the parity tests and bench.

Record layout follows the reference's ``radar_pointcloud_<n>.bin`` frames: five float32 per point in the
order x, y, z, intensity, doppler (/root/reference/src/iterative_closest_point.cpp:373-377).  The
registration path consumes the first four as a packed ``xyzw`` row (w = intensity).

Scene: scatterers on a ground plane, a handful of vertical walls and 10 % uniform clutter inside
x,y in [-80, 80] m (RADAR_RADIUS 80, /root/reference/src/radar_odometry.cpp:36), z in [-3, 3] m;
intensity ~ U(0, 40) dB; per-point noise sigma = 0.05 m.  A frame pair is two independent re-samples
of the same surfaces (so no exact duplicates => tie-free distances) related by a ground-truth SE(3).
"""
from __future__ import annotations

import numpy as np

EXTENT = 80.0
Z_LO, Z_HI = -3.0, 3.0
GROUND_Z = -1.5
NOISE = 0.05
SIGMA_VD = 0.125  # reve sigma_v_d, radar_odometry.cpp:607


def se3(yaw=0.0, pitch=0.0, roll=0.0, t=(0.0, 0.0, 0.0)) -> np.ndarray:
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    return T


def random_small_se3(rng, max_t=0.5, max_yaw_deg=3.0, max_rp_deg=0.5) -> np.ndarray:
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    t = d * rng.uniform(0, max_t)
    t[2] *= 0.2
    return se3(np.deg2rad(rng.uniform(-max_yaw_deg, max_yaw_deg)), np.deg2rad(rng.uniform(-max_rp_deg, max_rp_deg)),
               np.deg2rad(rng.uniform(-max_rp_deg, max_rp_deg)), t)


class Scene:
    """A fixed set of surfaces (ground + walls) that can be re-sampled any number of times."""

    def __init__(self, seed: int, extent: float = EXTENT, n_walls: int = 12):
        rng = np.random.default_rng(seed)
        self.extent = float(extent)
        # wall = segment (x0,y0)-(x1,y1), full height
        c = rng.uniform(-0.8 * extent, 0.8 * extent, size=(n_walls, 2))
        ang = rng.uniform(0, np.pi, size=n_walls)
        half = rng.uniform(0.1 * extent, 0.3 * extent, size=n_walls)
        d = np.stack([np.cos(ang), np.sin(ang)], 1) * half[:, None]
        self.w0, self.w1 = c - d, c + d
        self.wlen = 2 * half

    def sample(self, rng, n: int, centre=(0.0, 0.0), radius: float | None = None) -> np.ndarray:
        """n points [n,4] float32 (x,y,z,intensity) in the world frame; if radius is given the ground and
        clutter are drawn inside the square of half-size radius around centre (a sensor footprint)."""
        ext = self.extent if radius is None else float(radius)
        cx, cy = centre
        n_w = int(0.3 * n)
        n_c = int(0.1 * n)
        n_g = n - n_w - n_c
        g = np.empty((n_g, 3))
        g[:, 0] = rng.uniform(cx - ext, cx + ext, n_g)
        g[:, 1] = rng.uniform(cy - ext, cy + ext, n_g)
        g[:, 2] = GROUND_Z
        wi = rng.choice(len(self.wlen), size=n_w, p=self.wlen / self.wlen.sum())
        u = rng.uniform(0, 1, n_w)[:, None]
        w = np.empty((n_w, 3))
        w[:, :2] = self.w0[wi] * (1 - u) + self.w1[wi] * u
        w[:, 2] = rng.uniform(GROUND_Z, Z_HI, n_w)
        cl = np.empty((n_c, 3))
        cl[:, 0] = rng.uniform(cx - ext, cx + ext, n_c)
        cl[:, 1] = rng.uniform(cy - ext, cy + ext, n_c)
        cl[:, 2] = rng.uniform(Z_LO, Z_HI, n_c)
        p = np.concatenate([g, w, cl], 0)
        p += rng.normal(0, NOISE, p.shape)
        out = np.empty((n, 4), np.float32)
        out[:, :3] = p
        out[:, 3] = rng.uniform(0, 40, n)
        return out[rng.permutation(n)]


def apply(T: np.ndarray, pts: np.ndarray) -> np.ndarray:
    out = pts.copy()
    out[:, :3] = (pts[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    return out


def frame_pair(seed: int, n: int, m: int | None = None, extent: float = EXTENT, max_t=0.5, max_yaw_deg=3.0):
    """(src [n,4], tgt [m,4], T_gt) with tgt ~= T_gt * src surfaces (T_gt maps source frame -> target frame)."""
    m = n if m is None else m
    rng = np.random.default_rng(seed)
    sc = Scene(seed, extent)
    tgt = sc.sample(rng, m)
    src_w = sc.sample(rng, n)
    T_gt = random_small_se3(rng, max_t, max_yaw_deg)
    src = apply(np.linalg.inv(T_gt), src_w)
    return src, tgt, T_gt


def scan_to_map(seed: int, n: int, m: int, extent: float = EXTENT, scan_radius: float = 40.0):
    """(scan [n,4], map [m,4], T_gt): an accumulated map of the scene and one scan of its central area."""
    rng = np.random.default_rng(seed)
    sc = Scene(seed, extent)
    mp = sc.sample(rng, m)
    scan_w = sc.sample(rng, n, radius=scan_radius)
    T_gt = random_small_se3(rng)
    scan = apply(np.linalg.inv(T_gt), scan_w)
    return scan, mp, T_gt


def dense_map(seed: int, m: int, size=(400.0, 400.0, 20.0)) -> np.ndarray:
    """C5-style dense map: uniform volume density over a box centred on the origin (~6.25 pts/m^3 at 20 M)."""
    rng = np.random.default_rng(seed)
    out = np.empty((m, 4), np.float32)
    for a in range(3):
        out[:, a] = rng.uniform(-size[a] / 2, size[a] / 2, m)
    out[:, 3] = rng.uniform(0, 40, m)
    return out


def doppler(rng, pts_sensor: np.ndarray, v_ego, dynamic_frac: float = 0.1) -> np.ndarray:
    """Radial velocity per point for a sensor moving with v_ego: v_r = -u.v_ego + N(0, 0.125^2); a fraction of
    'dynamic' points gets U(-5, 5) m/s added. Returns [n] float32."""
    p = pts_sensor[:, :3].astype(np.float64)
    r = np.linalg.norm(p, axis=1)
    u = p / np.maximum(r, 1e-9)[:, None]
    vr = -(u @ np.asarray(v_ego, np.float64)) + rng.normal(0, SIGMA_VD, len(p))
    dyn = rng.uniform(size=len(p)) < dynamic_frac
    vr[dyn] += rng.uniform(-5, 5, int(dyn.sum()))
    return vr.astype(np.float32), dyn


def radar_frame_bin(pts_xyzi: np.ndarray, vr: np.ndarray) -> np.ndarray:
    """[n,5] float32 in the reference's .bin order x,y,z,intensity,doppler."""
    out = np.empty((len(pts_xyzi), 5), np.float32)
    out[:, :4] = pts_xyzi
    out[:, 4] = vr
    return out


def trajectory(seed: int, frames: int, step=0.4, max_yaw_deg=2.0):
    """Smooth 2-D trajectory: list of world poses T_w_s (<= 0.5 m, <= 3 deg per frame)."""
    rng = np.random.default_rng(seed)
    yaw, x, y = 0.0, 0.0, 0.0
    rate = 0.0
    poses = []
    for _ in range(frames):
        poses.append(se3(yaw, 0, 0, (x, y, 0.0)))
        rate = 0.95 * rate + 0.05 * rng.uniform(-max_yaw_deg, max_yaw_deg)
        yaw += np.deg2rad(rate)
        x += step * np.cos(yaw)
        y += step * np.sin(yaw)
    return poses
