"""Long-running differential fuzz: the CPU oracle (switched to the reference's operation order, tests/test_ref_sources.py)
against the reference's own sources on mini-Eigen (oracle/_ref/libplref2.so).  Needs /root/reference.  Everything must be
bit-identical; the script prints the number of mismatches per section.

    python tools/fuzz_reference_sources.py [scale]     # scale 1.0: ~6000 estimate_* problems, 60000 minimal scenes, 3000 cameras
Last full run (round 1): 0 mismatches in 6000 estimate_* problems (sizes 5..100, all four kinds and losses, PROSAC, degenerate
data), 42168 minimal scenes x 6 solver entry points (round 2, scale 0.5, with p3p_lambdatwist as a seventh solver entry point and mini-Eigen's coefficient-wise array():
0 mismatches in 3000 estimate_* problems, 21130 minimal scenes, 4500 camera calls) (planar, pure rotation, collinear, duplicated, noisy), 9000 camera calls.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import plo_py as P  # noqa: E402
import test_ref_sources as T  # noqa: E402
from poselib_b200 import problem_generator as G  # noqa: E402

CAMT = (G.FOCAL, G.FOCAL, 0.0, 0.0)


def both(f):
    a = f()
    with P.reference_sources():
        b = f()
    return a, b


def same(a, b):
    if isinstance(a, dict):
        return all(same(a[k], b[k]) for k in a if k != "counters")
    if isinstance(a, (tuple, list)):
        return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def fuzz_estimate(count):
    rng = np.random.default_rng(123)
    bad = 0
    for it in range(count):
        kind = ["relpose", "fundamental", "homography", "pnp"][it % 4]
        n = int(rng.choice([5, 6, 7, 8, 9, 12, 15, 25, 40, 100]))
        ratio = float(rng.choice([0.1, 0.3, 0.6, 0.9, 1.0]))
        mode = int(rng.integers(0, 5))
        if kind == "pnp":
            p = G.abspose_problem(max(n, 10), ratio, 1, it)
            a1, a2, thr, kw = p["x"][:n].copy(), p["X"][:n].copy(), float(rng.choice([2.0, 12.0])), dict(cam1=CAMT)
            if mode == 1:
                a2[:, 2] = a2[:, 2].mean()
            if mode == 2:
                a1[:n // 2], a2[:n // 2] = a1[0], a2[0]
        elif kind == "homography":
            p = G.homography_problem(max(n, 10), ratio, 4, it)
            a1, a2, thr, kw = p["x1"][:n].copy(), p["x2"][:n].copy(), float(rng.choice([0.5, 3.0])), {}
            if mode == 2:
                a1[:n // 2], a2[:n // 2] = a1[0], a2[0]
            if mode == 3:
                a1[:, 1] = a1[:, 0] * 0.5 + 3
        else:
            p = G.relpose_problem(max(n, 10), ratio, 2, it)
            a1, a2, thr = p["x1"][:n].copy(), p["x2"][:n].copy(), float(rng.choice([0.5, 3.0]))
            kw = dict(cam1=CAMT, cam2=CAMT) if kind == "relpose" else {}
            if mode == 1:
                a2 = a1 + rng.normal(0, 0.3, a1.shape)
            if mode == 2:
                a1[:n // 2], a2[:n // 2] = a1[0], a2[0]
            if mode == 3:
                a2 = a1 * 1.1
        ro = P.RansacOpt(max_iterations=int(rng.choice([50, 500, 3000])), min_iterations=int(rng.choice([10, 100])),
                         seed=int(rng.integers(0, 1 << 30)), progressive_sampling=bool(rng.integers(0, 2)))
        bo = P.BundleOpt(loss_type=str(rng.choice(["TRIVIAL", "TRUNCATED", "HUBER", "CAUCHY"])))
        a, b = both(lambda: P.estimate(kind, a1, a2, ro, bo, thr, **kw))
        bad += not same(a, b)
    return bad


def fuzz_minimal(count):
    rng = np.random.default_rng(5)
    unit = lambda a: a / np.linalg.norm(a, axis=1)[:, None]  # noqa: E731
    bad, used = {}, 0
    for it in range(count):
        mode = it % 6
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        R *= np.sign(np.linalg.det(R))
        t = rng.normal(size=3)
        X = np.c_[rng.uniform(-1, 1, (8, 2)), rng.uniform(2, 6, 8)]
        if mode == 1:
            X[:, 2] = 4.0
        if mode == 2:
            t = np.zeros(3) + 1e-9
        if mode == 3:
            X[2] = 0.5 * (X[0] + X[1])
        if mode == 4:
            X[1] = X[0]
        if mode == 5:
            X += rng.normal(0, 0.05, X.shape)
        Y = X @ R.T + t
        if (Y[:, 2] <= 0.1).any():
            continue
        used += 1
        x1, x2 = unit(X), unit(Y)
        if mode == 5:
            x2 = unit(Y + rng.normal(0, 0.01, Y.shape))
        for name, f in (("p3p", lambda: P.p3p(x2[:3], X[:3])), ("p3p_lambdatwist", lambda: P.p3p_lambdatwist(x2[:3], X[:3])),
                        ("relpose_5pt_E", lambda: P.relpose_5pt_E(x1[:5], x2[:5])),
                        ("relpose_5pt", lambda: P.relpose_5pt(x1[:5], x2[:5])), ("relpose_7pt", lambda: P.relpose_7pt(x1[:7], x2[:7])),
                        ("homography_4pt", lambda: P.homography_4pt(x1[:4], x2[:4])[1]),
                        ("essential_matrix_8pt", lambda: P.essential_matrix_8pt(x1, x2))):
            a, b = both(f)
            if not same(a, b):
                bad[name] = bad.get(name, 0) + 1
    return used, bad


def fuzz_cameras(count):
    rng = np.random.default_rng(11)
    bad = calls = 0
    for it in range(count):
        f = float(rng.uniform(200, 3000))
        cx, cy = rng.uniform(-200, 200, 2)
        k = rng.normal(0, 0.3 if it % 3 else 0.02, 2)
        pt = rng.normal(0, 0.01, 2)
        cam = [("SIMPLE_PINHOLE", [f, cx, cy]), ("PINHOLE", [f, f * rng.uniform(0.8, 1.2), cx, cy]),
               ("SIMPLE_RADIAL", [f, cx, cy, k[0]]), ("RADIAL", [f, cx, cy, k[0], k[1]]),
               ("OPENCV", [f, f * rng.uniform(0.8, 1.2), cx, cy, k[0], k[1], pt[0], pt[1]])][it % 5]
        X = np.c_[rng.uniform(-1.5, 1.5, (16, 2)), np.ones(16)]
        X[0, :2] = 0
        X /= np.linalg.norm(X, axis=1)[:, None]
        a, b = both(lambda: P.camera_project_with_jac(cam, X))
        calls += 1
        if not same(a, b):
            bad += 1
            continue
        xp = a[2] + rng.normal(0, 50, (16, 2)) * (it % 2)
        for fn in (P.camera_unproject_with_jac, P.camera_unproject2):
            a2, b2 = both(lambda: fn(cam, xp))
            calls += 1
            bad += not same(a2, b2)
    return calls, bad


if __name__ == "__main__":
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    P.set_reference_order(True, T._parse_reference_det_terms())
    try:
        n = int(6000 * scale)
        print("estimate_*:", n, "problems, mismatches", fuzz_estimate(n))
        used, bad = fuzz_minimal(int(60000 * scale))
        print("minimal solvers:", used, "scenes x 7 entry points, mismatches", bad)
        calls, bad = fuzz_cameras(int(3000 * scale))
        print("camera models:", calls, "calls, mismatches", bad)
    finally:
        P.set_reference_order(False)
