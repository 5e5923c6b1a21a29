"""Synthetic LO-RANSAC problem instances for the BASELINE.json configs.

Extends the distributions of the reference's minimal-instance generator
(/root/reference/benchmark/problem_generator.cc:338-405 abs. pose, :537-639 rel. pose, :641-743 homography;
FoV 75 deg from solver_benchmark.cc:351; depth U(0.1,10) problem_generator.h:107-108; GT pose = random unit
quaternion + t in U(-1,1)^3, :320-335) with what the RANSAC-level configs need and the reference generator
lacks: N >> sample size, outliers, pixel noise through a virtual PINHOLE camera (f = 1000 px, pp = 0) and a
PROSAC quality ordering.  Everything is a pure function of (config, problem_idx) via numpy's MT19937.
"""
import math

import numpy as np

FOCAL = 1000.0
FOV_DEG = 75.0
FOV_SCALE = math.tan(FOV_DEG / 2.0 * math.pi / 180.0)
NOISE_PX = 0.5


def _rng(config_id, problem_idx):
    return np.random.Generator(np.random.MT19937(0xB2000000 + config_id * 1000 + problem_idx))


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return quat_to_rotmat(q), q


def quat_to_rotmat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _bearings(rng, n):
    uv = rng.uniform(-FOV_SCALE, FOV_SCALE, size=(n, 2))
    x = np.concatenate([uv, np.ones((n, 1))], axis=1)
    return x / np.linalg.norm(x, axis=1, keepdims=True), uv


def abspose_problem(n, inlier_ratio, config_id=1, problem_idx=0):
    """2D-3D correspondences.  Returns pixel 2D points (f=1000, pp=0), 3D points, GT pose, inlier flags."""
    rng = _rng(config_id, problem_idx)
    R, q = _random_rotation(rng)
    t = rng.uniform(-1, 1, size=3)
    x, uv = _bearings(rng, n)
    depth = rng.uniform(0.1, 10.0, size=n)
    Xc = x * depth[:, None]
    X = (Xc - t) @ R  # R^T (Xc - t), row form
    px = FOCAL * uv + rng.normal(0, NOISE_PX, size=(n, 2))
    n_in = int(round(n * inlier_ratio))
    is_inl = np.zeros(n, dtype=bool)
    is_inl[rng.permutation(n)[:n_in]] = True
    out = ~is_inl
    px[out] = FOCAL * rng.uniform(-FOV_SCALE, FOV_SCALE, size=(out.sum(), 2))
    return {"x": px, "X": X, "q_gt": q, "t_gt": t, "R_gt": R, "is_inlier": is_inl,
            "camera": (FOCAL, FOCAL, 0.0, 0.0)}


def relpose_problem(n, inlier_ratio, config_id=2, problem_idx=0, prosac_sorted=False):
    """2D-2D correspondences in pixels for two PINHOLE f=1000 cameras; |t| = 1."""
    rng = _rng(config_id, problem_idx)
    # Additions to the minimal generator (which checks nothing): the pose is redrawn until the two views overlap,
    # and only points in front of and inside the field of view of camera 2 are kept.
    while True:
        R, q = _random_rotation(rng)
        t = rng.uniform(-1, 1, size=3)
        t /= np.linalg.norm(t)
        x1, uv = _bearings(rng, 256)
        X2 = (x1 * rng.uniform(0.1, 10.0, size=256)[:, None]) @ R.T + t
        vis = (X2[:, 2] > 0.05) & (np.abs(X2[:, :2] / np.maximum(X2[:, 2:3], 1e-9)) <= FOV_SCALE).all(axis=1)
        if vis.mean() >= 0.2:
            break
    uv1 = np.zeros((n, 2))
    uv2 = np.zeros((n, 2))
    filled = 0
    while filled < n:
        m = max(2 * (n - filled), 64)
        x1, uv = _bearings(rng, m)
        depth = rng.uniform(0.1, 10.0, size=m)
        X2 = (x1 * depth[:, None]) @ R.T + t
        h = X2[:, :2] / np.maximum(X2[:, 2:3], 1e-9)
        ok = (X2[:, 2] > 0.05) & (np.abs(h) <= FOV_SCALE).all(axis=1)
        k = min(int(ok.sum()), n - filled)
        uv1[filled:filled + k] = uv[ok][:k]
        uv2[filled:filled + k] = h[ok][:k]
        filled += k
    px1 = FOCAL * uv1 + rng.normal(0, NOISE_PX, size=(n, 2))
    px2 = FOCAL * uv2 + rng.normal(0, NOISE_PX, size=(n, 2))
    n_in = int(round(n * inlier_ratio))
    is_inl = np.zeros(n, dtype=bool)
    is_inl[rng.permutation(n)[:n_in]] = True
    out = ~is_inl
    px2[out] = FOCAL * rng.uniform(-FOV_SCALE, FOV_SCALE, size=(out.sum(), 2))
    if prosac_sorted:  # synthetic match quality: inliers ~U(0.4,1), outliers ~U(0,0.7); sort descending
        qual = np.where(is_inl, rng.uniform(0.4, 1.0, size=n), rng.uniform(0.0, 0.7, size=n))
        order = np.argsort(-qual, kind="stable")
        px1, px2, is_inl = px1[order], px2[order], is_inl[order]
    return {"x1": px1, "x2": px2, "q_gt": q, "t_gt": t, "R_gt": R, "is_inlier": is_inl,
            "camera": (FOCAL, FOCAL, 0.0, 0.0)}


def homography_problem(n, inlier_ratio, config_id=4, problem_idx=0):
    """Points on a random plane seen by two cameras (problem_generator.cc:676-716), pixels at f=1000."""
    rng = _rng(config_id, problem_idx)
    while True:
        R, q = _random_rotation(rng)
        t = rng.uniform(-1, 1, size=3)
        t /= np.linalg.norm(t)
        nrm = rng.normal(size=3)
        nrm /= np.linalg.norm(nrm)
        d_center = rng.uniform(0.1, 10.0)
        alpha = d_center / nrm[2]
        H_gt = alpha * R + np.outer(t, nrm)
        uv1 = np.zeros((n, 2))
        uv2 = np.zeros((n, 2))
        filled = 0
        tries = 0
        while filled < n and tries < 200:
            tries += 1
            m = max(n - filled, 64)
            x1, uv = _bearings(rng, m)
            lam = alpha / (x1 @ nrm)
            X2 = (x1 * lam[:, None]) @ R.T + t
            h = X2[:, :2] / X2[:, 2:3]
            ok = (lam > 0) & (X2[:, 2] > 0) & (np.abs(h) <= FOV_SCALE).all(axis=1)
            k = min(int(ok.sum()), n - filled)
            uv1[filled:filled + k] = uv[ok][:k]
            uv2[filled:filled + k] = h[ok][:k]
            filled += k
        if filled == n:
            break
    px1 = FOCAL * uv1 + rng.normal(0, NOISE_PX, size=(n, 2))
    px2 = FOCAL * uv2 + rng.normal(0, NOISE_PX, size=(n, 2))
    n_in = int(round(n * inlier_ratio))
    is_inl = np.zeros(n, dtype=bool)
    is_inl[rng.permutation(n)[:n_in]] = True
    out = ~is_inl
    px2[out] = FOCAL * rng.uniform(-FOV_SCALE, FOV_SCALE, size=(out.sum(), 2))
    return {"x1": px1, "x2": px2, "H_gt": H_gt, "is_inlier": is_inl, "camera": (FOCAL, FOCAL, 0.0, 0.0)}


# ---- the five BASELINE.json configs ------------------------------------------------------------------
def config_c1(problem_idx=0, n=200):
    """p3p absolute pose, 200 corrs, 50 % inliers, max_error 12 px, exactly 1000 iterations."""
    p = abspose_problem(n, 0.5, 1, problem_idx)
    p["ransac"] = dict(max_iterations=1000, min_iterations=1000)
    p["max_error"] = 12.0
    return p


def config_c2(problem_idx=0, n=10000):
    """relpose_5pt, 10 000 corrs, 30 % inliers, max_error 1 px, max 100 000 its (headline config)."""
    p = relpose_problem(n, 0.3, 2, problem_idx)
    p["ransac"] = dict(max_iterations=100000, min_iterations=1000)
    p["max_error"] = 1.0
    return p


def config_c3(problem_idx=0, n=5000):
    """relpose_7pt + real_focal_check + PROSAC, 5 000 corrs, 20 % inliers."""
    p = relpose_problem(n, 0.2, 3, problem_idx, prosac_sorted=True)
    p["ransac"] = dict(max_iterations=100000, min_iterations=1000, progressive_sampling=True,
                       max_prosac_iterations=100000)
    p["max_error"] = 1.0
    p["real_focal_check"] = True
    return p


def config_c4(problem_idx=0, n=20000):
    """homography_4pt, 20 000 corrs, 60 % inliers, LO refit with TRUNCATED loss."""
    p = homography_problem(n, 0.6, 4, problem_idx)
    p["ransac"] = dict(max_iterations=100000, min_iterations=1000)
    p["max_error"] = 1.0
    return p


def config_c5(count=4096, first=0):
    """Batch of independent problems alternating C1-type (even idx) and C2-type (odd idx)."""
    out = []
    for i in range(first, first + count):
        if i % 2 == 0:
            p = abspose_problem(200, 0.5, 5, i)
            p.update(kind="pnp", ransac=dict(max_iterations=1000, min_iterations=1000), max_error=12.0)
        else:
            p = relpose_problem(10000, 0.3, 5, i)
            p.update(kind="relpose", ransac=dict(max_iterations=100000, min_iterations=1000), max_error=1.0)
        out.append(p)
    return out


# ---- minimal noise-free instances (solver validity KATs, solver_benchmark.cc:27-45) -------------------
def minimal_abspose(problem_idx):
    rng = _rng(11, problem_idx)
    R, q = _random_rotation(rng)
    t = rng.uniform(-1, 1, size=3)
    x, _ = _bearings(rng, 3)
    X = (x * rng.uniform(0.1, 10.0, size=3)[:, None] - t) @ R
    return x, X, R, t


def minimal_relpose(problem_idx, npts):
    rng = _rng(12, problem_idx)
    R, q = _random_rotation(rng)
    t = rng.uniform(-1, 1, size=3)
    t /= np.linalg.norm(t)
    x1, _ = _bearings(rng, npts)
    X2 = (x1 * rng.uniform(0.1, 10.0, size=npts)[:, None]) @ R.T + t
    x2 = X2 / np.linalg.norm(X2, axis=1, keepdims=True)
    return x1, x2, R, t


def minimal_homography(problem_idx):
    rng = _rng(14, problem_idx)
    while True:
        R, q = _random_rotation(rng)
        t = rng.uniform(-1, 1, size=3)
        t /= np.linalg.norm(t)
        nrm = rng.normal(size=3)
        nrm /= np.linalg.norm(nrm)
        alpha = rng.uniform(0.1, 10.0) / nrm[2]
        H = alpha * R + np.outer(t, nrm)
        x1, _ = _bearings(rng, 64)
        lam = alpha / (x1 @ nrm)
        X2 = (x1 * lam[:, None]) @ R.T + t
        h = X2[:, :2] / X2[:, 2:3]
        ok = (lam > 0) & (X2[:, 2] > 0) & (np.abs(h) <= FOV_SCALE).all(axis=1)
        if ok.sum() >= 4:
            x1 = x1[ok][:4]
            X2 = X2[ok][:4]
            return x1, X2 / np.linalg.norm(X2, axis=1, keepdims=True), H
