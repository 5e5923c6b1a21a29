"""BASELINE config 5: a batch of independent problems alternating C1-type (p3p, 200 corrs) and C2-type (5pt, 10 000
corrs), own data seed each, through ONE plb_ransac_batch call (host buffers in, results out).  Prints problems/s,
hypotheses/s and the wall time of the call; `--check K` compares the first K results with the CPU oracle."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poselib_b200 import cabi, problem_generator as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--count", type=int, default=4096)
ap.add_argument("--streams", type=int, default=12)
ap.add_argument("--mode", default="fast")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--check", type=int, default=0)
a = ap.parse_args()

t0 = time.perf_counter()
raw = G.config_c5(a.count)
probs = []
for p in raw:
    if p["kind"] == "pnp":
        A, B, me = np.ascontiguousarray(p["x"] / G.FOCAL), np.ascontiguousarray(p["X"]), p["max_error"] / G.FOCAL
    else:
        A, B, me = np.ascontiguousarray(p["x1"] / G.FOCAL), np.ascontiguousarray(p["x2"] / G.FOCAL), p["max_error"] / G.FOCAL
    probs.append(dict(kind=p["kind"], a=A, b=B, ransac=cabi.RansacOpt(**p["ransac"]), max_error=me))
t_gen = time.perf_counter() - t0
cabi.set_mode(a.mode)
best = None
for rep in range(a.reps):
    t0 = time.perf_counter()
    res = cabi.ransac_batch(probs, streams=a.streams)
    dt = time.perf_counter() - t0
    hyp = sum(r["counters"]["hypotheses"] for r in res)
    cor = sum(r["counters"]["scored_corrs"] for r in res)
    line = {"config": "C5", "count": a.count, "mode": a.mode, "streams": a.streams, "rep": rep, "seconds": dt,
            "problems_per_s": a.count / dt, "hypotheses_per_s": hyp / dt, "scored_corrs_per_s": cor / dt,
            "inliers_pnp_mean": float(np.mean([r["stats"]["num_inliers"] for r, p in zip(res, probs) if p["kind"] == "pnp"])),
            "inliers_relpose_mean": float(np.mean([r["stats"]["num_inliers"] for r, p in zip(res, probs) if p["kind"] == "relpose"])),
            "failed": int(sum(r["status"] != 0 for r in res)), "generate_seconds": t_gen}
    print(json.dumps(line), flush=True)
if a.check:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import plo_py as P
    bad = 0
    for r, p, q in list(zip(res, probs, raw))[:a.check]:
        o = P.ransac(p["kind"], p["a"], p["b"], P.RansacOpt(**q["ransac"]), p["max_error"])
        same = all(r["stats"][k] == o["stats"][k] for k in ("iterations", "refinements", "num_inliers")) and \
            np.array_equal(r["inliers"], o["inliers"])
        bad += not same
    print(json.dumps({"checked": a.check, "mismatches": bad}))
