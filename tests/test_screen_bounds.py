"""CPU tests of the fp32 screening arithmetic and its error bounds (poselib_b200/csrc/screen_math.cuh, compiled for the
host by tests/screen_bounds_host.cc): the bracket [count32 +- border] x [score32 +- err] that k_screen hands to the
candidate selection must contain the fp64 record for EVERY model — good, perturbed, random and degenerate ones — and the
streaming test must never rule out a correspondence whose fp64 residual is under the threshold.  Data: the synthetic
problems of the bench configs, the same in pixel units and with large offsets (where fp32 is hopeless and everything
must come out as 'uncertain' rather than wrong), and adversarial point sets whose residuals sit within 1e-3 .. 1e-9
relative of the threshold."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from poselib_b200 import problem_generator as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "screen_bounds_host.cc")
HDR = os.path.join(ROOT, "poselib_b200", "csrc", "screen_math.cuh")
LIB = os.path.join(ROOT, "tests", "_screen_host.so")
_P = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", SRC, "-o", LIB])
    return C.CDLL(LIB)


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_P)


def check_sampson(lib, kind, pts, M, qt, sq_thr):
    pts, pp = _d(pts)
    M, mp = _d(M)
    qt, qp = _d(qt if qt is not None else np.zeros(7))
    out = np.zeros(8)
    lib.scr_check_sampson(C.c_int(kind), C.c_int(pts.shape[1]), pp, mp, qp, C.c_double(sq_thr), out.ctypes.data_as(_P))
    return dict(viol=int(out[0]), c64=int(out[1]), c32=int(out[2]), border=int(out[3]), s64=out[4], s32=out[5], err=out[6],
                maybe=int(out[7]))


def check_transfer(lib, kind, pts, M, sq_thr):
    pts, pp = _d(pts)
    M, mp = _d(M)
    out = np.zeros(8)
    lib.scr_check_transfer(C.c_int(kind), C.c_int(pts.shape[1]), pp, mp, C.c_double(sq_thr), out.ctypes.data_as(_P))
    return dict(viol=int(out[0]), c64=int(out[1]), c32=int(out[2]), border=int(out[3]), s64=out[4], s32=out[5], err=out[6],
                maybe=int(out[7]))


def essential(R, t):
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    return tx @ R


def rot_to_quat(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    if w > 1e-6:
        return np.array([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])
    return np.array([0.0, 1.0, 0.0, 0.0])


def small_rot(rng, s):
    w = rng.normal(0, s, 3)
    th = np.linalg.norm(w)
    if th == 0:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def relpose_models(rng, p):
    """(R, t) close to and far from the ground truth: the screening pass sees all of them"""
    out = [(p["R_gt"], p["t_gt"])]
    for s in (1e-4, 1e-3, 1e-2, 0.1):
        for _ in range(3):
            t = p["t_gt"] + rng.normal(0, s, 3)
            out.append((small_rot(rng, s) @ p["R_gt"], t / np.linalg.norm(t)))
    for _ in range(6):
        t = rng.normal(0, 1, 3)
        out.append((small_rot(rng, 2.0), t / np.linalg.norm(t)))
    return out


@pytest.mark.parametrize("idx", range(6))
def test_relpose_and_fundamental_records_bracket_fp64(lib, idx):
    rng = np.random.default_rng(100 + idx)
    p = G.relpose_problem(4000, 0.4, config_id=31, problem_idx=idx)
    stats = []
    for scale, thr_px in ((1.0 / G.FOCAL, 1.0), (1.0 / G.FOCAL, 0.05), (1.0, 1.0), (1.0 / G.FOCAL, 12.0)):
        pts = np.vstack([p["x1"].T, p["x2"].T]) * scale  # normalised coordinates, or raw pixels (scale 1)
        for R, t in relpose_models(rng, p):
            E = essential(R, t)
            if scale == 1.0:  # pixel units: F = K^-T E K^-1
                Kinv = np.diag([1 / G.FOCAL, 1 / G.FOCAL, 1.0])
                E = Kinv.T @ E @ Kinv
                E = E / np.linalg.norm(E)
            sq_thr = (thr_px * (scale if scale != 1.0 else 1.0)) ** 2
            qt = np.r_[rot_to_quat(R), t]
            r1 = check_sampson(lib, 1 if scale != 1.0 else 2, pts, E.reshape(-1), qt, sq_thr)
            r2 = check_sampson(lib, 2, pts, E.reshape(-1), None, sq_thr)
            assert r1["viol"] == 0 and r2["viol"] == 0, (scale, thr_px, r1, r2)
            stats.append((scale, r2["c64"], r2["border"], r2["err"] / max(r2["s64"], 1e-300)))
    # tightness on well-scaled data: at most a few percent of the inliers are uncertain for the ground-truth model
    gt = [s for s in stats if s[0] != 1.0 and s[1] > 1000]
    assert gt and all(b <= 0.05 * c + 5 for _, c, b, _ in gt), gt[:4]


@pytest.mark.parametrize("idx", range(4))
def test_homography_records_bracket_fp64(lib, idx):
    rng = np.random.default_rng(200 + idx)
    p = G.homography_problem(4000, 0.6, config_id=32, problem_idx=idx)
    for scale, thr in ((1.0 / G.FOCAL, 1.0 / G.FOCAL), (1.0, 1.0), (1.0, 0.1), (1.0 / G.FOCAL, 1e-5)):
        pts = np.vstack([p["x1"].T, p["x2"].T]) * scale
        K = np.diag([G.FOCAL, G.FOCAL, 1.0]) if scale == 1.0 else np.eye(3)
        H0 = K @ p["H_gt"] @ np.linalg.inv(K)
        H0 = H0 / np.linalg.norm(H0)
        models = [H0] + [H0 + rng.normal(0, s, (3, 3)) for s in (1e-6, 1e-4, 1e-2, 1.0) for _ in range(3)]
        models.append(np.array([[1, 0, 0], [0, 1, 0], [0.7, 0.1, 1e-3]]))  # vanishing line through the image
        for H in models:
            r = check_transfer(lib, 3, pts, (H / np.linalg.norm(H)).reshape(-1), thr * thr)
            assert r["viol"] == 0, (scale, thr, r)


@pytest.mark.parametrize("idx", range(4))
def test_pnp_records_bracket_fp64(lib, idx):
    rng = np.random.default_rng(300 + idx)
    p = G.abspose_problem(3000, 0.5, config_id=33, problem_idx=idx)
    for offset, thr in ((0.0, 12.0 / G.FOCAL), (0.0, 0.5 / G.FOCAL), (1e3, 12.0 / G.FOCAL), (1e5, 1.0 / G.FOCAL)):
        X = p["X"] + offset  # world frame far from the origin: fp32 loses the scene, the bounds must say so
        x = p["x"] / G.FOCAL
        pts = np.vstack([x.T, X.T])
        R, t = p["R_gt"], p["t_gt"] - p["R_gt"] @ np.full(3, offset)
        models = [(R, t)]
        for s in (1e-4, 1e-2, 0.3):
            for _ in range(3):
                Rp = small_rot(rng, s) @ R
                models.append((Rp, t + rng.normal(0, s, 3) - (Rp - R) @ np.full(3, offset)))
        models.append((small_rot(rng, 2.0), rng.normal(0, 3, 3)))
        for Rm, tm in models:
            r = check_transfer(lib, 0, pts, np.hstack([Rm, tm[:, None]]).reshape(-1), thr * thr)
            assert r["viol"] == 0, (offset, thr, r)


def test_residuals_on_the_threshold(lib):
    """Adversarial: thousands of correspondences whose residual is within 1e-3 ... 1e-9 (relative) of the threshold — the
    case the round-1 heuristic margins (4 + 1 % of the count) had no answer to.  Every one of them must be either
    decided like fp64 or flagged uncertain, and the counts must bracket."""
    rng = np.random.default_rng(7)
    p = G.relpose_problem(2000, 1.0, config_id=34, problem_idx=0)
    E = essential(p["R_gt"], p["t_gt"])
    x1 = p["x1"] / G.FOCAL
    x2 = p["x2"] / G.FOCAL
    thr = 1.0 / G.FOCAL
    # move x2 along the epipolar line normal until the Sampson residual equals thr (1 + eps)
    n = len(x1)
    h1 = np.c_[x1, np.ones(n)]
    for eps_mag in (1e-3, 1e-5, 1e-7, 1e-9):
        l = h1 @ E.T  # epipolar lines in image 2
        nrm = l[:, :2] / np.linalg.norm(l[:, :2], axis=1, keepdims=True)
        x2p = x2.copy()
        target = thr * (1 + rng.choice([-1, 1], n) * eps_mag * rng.uniform(0.1, 1, n))
        for _ in range(60):  # fixed-point iteration on the signed Sampson distance
            h2 = np.c_[x2p, np.ones(n)]
            Cc = np.einsum("ij,ij->i", h2, l)
            Ex2 = h2 @ E
            den = np.sqrt(l[:, 0] ** 2 + l[:, 1] ** 2 + Ex2[:, 0] ** 2 + Ex2[:, 1] ** 2)
            r = Cc / den
            x2p += ((np.sign(r) * target - r) * (den / np.linalg.norm(l[:, :2], axis=1)))[:, None] * nrm
        pts = np.vstack([x1.T, x2p.T])
        r = check_sampson(lib, 2, pts, E.reshape(-1), None, thr * thr)
        assert r["viol"] == 0, (eps_mag, r)
        if eps_mag <= 1e-7:  # far inside fp32 resolution: nearly everything must be reported as uncertain
            assert r["border"] > 0.9 * n, (eps_mag, r)
        qt = np.r_[rot_to_quat(p["R_gt"]), p["t_gt"]]
        r = check_sampson(lib, 1, pts, E.reshape(-1), qt, thr * thr)
        assert r["viol"] == 0, (eps_mag, r)


def test_degenerate_models_and_inputs(lib):
    rng = np.random.default_rng(9)
    pts = rng.uniform(-0.8, 0.8, (4, 3000))
    for M in (np.zeros(9), np.r_[np.zeros(8), 1.0], np.r_[1e-30, np.zeros(8)], np.full(9, 1e18), rng.normal(0, 1e-12, 9)):
        for thr in (1e-6, 1e-12, 0.0, 1e6):
            assert check_sampson(lib, 2, pts, M, None, thr)["viol"] == 0
            assert check_transfer(lib, 3, pts, M, thr)["viol"] == 0
    bad = pts.copy()
    bad[1, 5] = np.nan
    bad[2, 9] = np.inf
    E = rng.normal(0, 1, 9)
    assert check_sampson(lib, 2, bad, E, None, 1e-6)["viol"] == 0
    assert check_transfer(lib, 3, bad, E, 1e-6)["viol"] == 0
    # a pose whose translation dwarfs the scene / a non-unit quaternion: the cheirality bound scales with both
    p = G.relpose_problem(2000, 0.5, config_id=35, problem_idx=1)
    pts = np.vstack([p["x1"].T, p["x2"].T]) / G.FOCAL
    for q_scale, t_scale in ((1.0, 1e3), (3.0, 1.0), (1.0, 1e-6), (0.1, 10.0)):
        qt = np.r_[rot_to_quat(p["R_gt"]) * q_scale, p["t_gt"] * t_scale]
        assert check_sampson(lib, 1, pts, essential(p["R_gt"], p["t_gt"]).reshape(-1), qt, 1e-6)["viol"] == 0
