"""Seeded synthetic depth scenes.  The reference's umbrella PNGs are downloaded at run time
(download_data.sh) and are not available offline, so the bench and tests use these surrogates.

Depth images are u16 millimetres, 0 = invalid, pinhole intrinsics K = (fx, fy, cx, cy) as kinfu.cpp:23.
"""
from __future__ import annotations

import numpy as np

DEFAULT_K = (570.342, 570.342, 320.0, 240.0)        # default_params_dynamicfusion, kinfu.cpp:23


def scaled_K(cols: int, rows: int):
    s = cols / 640.0
    return (DEFAULT_K[0] * s, DEFAULT_K[1] * s, cols / 2.0, rows / 2.0)


def volume_pose(size: float):
    """Affine3f().translate(-size/2, -size/2, 0.5), kinfu.cpp:27 -> (R, t)"""
    return np.eye(3, dtype=np.float32), np.array([-size / 2, -size / 2, 0.5], np.float32)


def _rays(cols, rows, K):
    fx, fy, cx, cy = K
    u, v = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    return (u - cx) / fx, (v - cy) / fy


def _finish(z, rng, noise_mm, dropout):
    """z (metres, nan = miss) -> u16 mm with uniform noise and random dropouts"""
    mm = z * 1000.0
    if noise_mm > 0:
        mm = mm + rng.uniform(-noise_mm, noise_mm, size=mm.shape)
    mm = np.where(np.isfinite(mm), mm, 0.0)
    if dropout > 0:
        mm = np.where(rng.random(mm.shape) < dropout, 0.0, mm)
    return np.clip(np.rint(mm), 0, 65535).astype(np.uint16)


def sphere_wall_depth(cols=640, rows=480, K=DEFAULT_K, seed=0, radius=0.25, centre=(0.0, 0.0, 1.0), wall_z=1.4,
                      noise_mm=1.0, dropout=0.02):
    """Config C1: sphere in front of a wall, camera at the origin looking down +z."""
    rng = np.random.default_rng(seed)
    xl, yl = _rays(cols, rows, K)
    # ray p = s*(xl, yl, 1); |p - c|^2 = r^2
    cx_, cy_, cz_ = centre
    a = xl * xl + yl * yl + 1.0
    b = -2.0 * (xl * cx_ + yl * cy_ + cz_)
    c = cx_ * cx_ + cy_ * cy_ + cz_ * cz_ - radius * radius
    disc = b * b - 4 * a * c
    s = np.where(disc >= 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.nan)
    z = np.where(np.isfinite(s), s, wall_z)
    return _finish(z, rng, noise_mm, dropout)


def camera_drift(t: int):
    """rigid drift of C2: 1 mm + 0.1 degree per frame (about the y axis) -> camera-to-world (R, t)"""
    ang = np.deg2rad(0.1) * t
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float64)
    tr = np.array([0.001 * t, 0.0, 0.0], np.float64)
    return R, tr


def umbrella_depth(t: int, cols=640, rows=480, K=DEFAULT_K, seed=0, wall_z=1.3, noise_mm=0.5, dropout=0.01, drift=True, shape_t=None):
    """Config C2 frame t: breathing paraboloid cap z = 0.9 + a(t)(x^2+y^2), radius 0.3 m, a(t) = 0.6+0.4 sin(2 pi t/50),
    in front of a wall; the camera drifts rigidly (camera_drift)."""
    rng = np.random.default_rng(seed * 100003 + t)
    xl, yl = _rays(cols, rows, K)
    R, tr = camera_drift(t) if drift else (np.eye(3), np.zeros(3))
    d = np.stack([xl, yl, np.ones_like(xl)], -1) @ R.T            # world-frame ray directions (unnormalised, z_cam = 1)
    o = tr
    a_t = 0.6 + 0.4 * np.sin(2 * np.pi * (t if shape_t is None else shape_t) / 50.0)
    # world point p = o + s*d ;  p.z = 0.9 + a (p.x^2 + p.y^2)
    A = a_t * (d[..., 0] ** 2 + d[..., 1] ** 2)
    B = 2 * a_t * (o[0] * d[..., 0] + o[1] * d[..., 1]) - d[..., 2]
    Cc = a_t * (o[0] ** 2 + o[1] ** 2) + 0.9 - o[2]
    disc = B * B - 4 * A * Cc
    with np.errstate(divide="ignore", invalid="ignore"):
        s_par = np.where(np.abs(A) > 1e-12, (-B - np.sqrt(np.maximum(disc, 0))) / (2 * A), -Cc / B)
    px = o[0] + s_par * d[..., 0]
    py = o[1] + s_par * d[..., 1]
    hit = (disc >= 0) & (s_par > 0) & (px * px + py * py <= 0.3 * 0.3)
    # tilted wall n.(p - p0) = 0 and a small static off-axis sphere: they break the cap's rotational symmetry so that
    # all six pose degrees of freedom are observable by ICP
    n = np.array([-0.10, -0.05, 1.0])
    p0 = np.array([0.0, 0.0, wall_z])
    s_wall = ((p0 - o) @ n) / (d @ n)
    sc, sr = np.array([0.27, -0.17, 1.02]), 0.07
    oc = o - sc
    qa = (d * d).sum(-1)
    qb = 2.0 * (d @ oc)
    qc = oc @ oc - sr * sr
    qd = qb * qb - 4 * qa * qc
    s_sph = np.where(qd >= 0, (-qb - np.sqrt(np.maximum(qd, 0))) / (2 * qa), np.inf)
    s_sph = np.where(s_sph > 0, s_sph, np.inf)
    s = np.where(hit, s_par, s_wall)            # camera-frame depth == s because the camera-frame ray has z = 1
    s = np.minimum(s, s_sph)
    return _finish(s, rng, noise_mm, dropout)


def make_pose_inverse(R, t):
    Ri = np.asarray(R, np.float64).T
    return Ri.astype(np.float32), (-Ri @ np.asarray(t, np.float64)).astype(np.float32)


def compose(A, B):
    """(RA,tA) * (RB,tB) in float32, row-dot-column order"""
    RA, tA = np.asarray(A[0], np.float32), np.asarray(A[1], np.float32)
    RB, tB = np.asarray(B[0], np.float32), np.asarray(B[1], np.float32)
    return (RA @ RB).astype(np.float32), (RA @ tB + tA).astype(np.float32)
