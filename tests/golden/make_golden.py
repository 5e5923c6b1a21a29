"""Regenerates tests/golden/oracle_regression.json.

These are REGRESSION fixtures of this repository's CPU oracle (oracle/), not outputs of the reference: PoseLib as a whole
cannot be built in this image (no Eigen); only its Eigen-free parts are (oracle/_ref).
They pin the oracle's behaviour between rounds: any change of its arithmetic or control flow shows up on the CPU, and
the GPU parity tests compare the CUDA path with the very same expectations.

    python tests/golden/make_golden.py        # rewrites the JSON next to this script
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import plo_py as P  # noqa: E402
from poselib_b200 import problem_generator as G  # noqa: E402

CASES = [  # (name, kind, generator call, ransac options, max_error in px)
    ("pnp_200", "pnp", lambda: G.abspose_problem(200, 0.5, 61, 0), dict(max_iterations=400, min_iterations=100, seed=1), 12.0),
    ("relpose_300", "relpose", lambda: G.relpose_problem(300, 0.4, 62, 0), dict(max_iterations=800, min_iterations=100, seed=2), 1.5),
    ("fundamental_300_prosac_rfc", "fundamental", lambda: G.relpose_problem(300, 0.5, 63, 0, prosac_sorted=True),
     dict(max_iterations=800, min_iterations=100, seed=3, progressive_sampling=True, max_prosac_iterations=500), 1.5),
    ("homography_300", "homography", lambda: G.homography_problem(300, 0.5, 64, 0), dict(max_iterations=400, min_iterations=50, seed=4), 1.5),
]


def hexlist(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).reshape(-1)]


def run_case(kind, p, kw, me_px):
    if kind == "pnp":
        a, b = p["x"] / G.FOCAL, p["X"]
    else:
        a, b = p["x1"] / G.FOCAL, p["x2"] / G.FOCAL
    r = P.ransac(kind, a, b, P.RansacOpt(**kw), me_px / G.FOCAL, rfc=(kind == "fundamental"))
    return a, b, r


def main():
    out = {"_about": "oracle regression fixtures (NOT reference outputs); regenerate with tests/golden/make_golden.py", "cases": {}}
    for name, kind, gen, kw, me in CASES:
        p = gen()
        a, b, r = run_case(kind, p, kw, me)
        k = {"pnp": 3, "relpose": 5, "fundamental": 7, "homography": 4}[kind]
        out["cases"][name] = {
            "kind": kind, "ransac": kw, "max_error_px": me,
            "first_samples": P.sample_table(len(a), k, P.RansacOpt(**kw), 8).tolist(),
            "stats": {q: r["stats"][q] for q in ("iterations", "refinements", "num_inliers")},
            "model_score": float(r["stats"]["model_score"]).hex(),
            "counters": {q: r["counters"][q] for q in ("samples", "hypotheses", "lo_calls")},
            "inliers": "".join("1" if v else "0" for v in r["inliers"]),
            "model": hexlist(r["model"]),
        }
    json.dump(out, open(os.path.join(HERE, "oracle_regression.json"), "w"), indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
