"""Runs one BASELINE config problem through the C-ABI a few times and prints stats/counters/timing (profiling aid)."""
import argparse
import os
import sys
import time


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poselib_b200 import cabi, problem_generator as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--mode", default="exact")
ap.add_argument("--idx", type=int, default=0)
a = ap.parse_args()
cabi.set_mode(a.mode)
if a.config == "c1":
    p = G.config_c1(a.idx); kind, A, B, me = "pnp", p["x"] / G.FOCAL, p["X"], 12.0 / G.FOCAL
elif a.config == "c2":
    p = G.config_c2(a.idx); kind, A, B, me = "relpose", p["x1"] / G.FOCAL, p["x2"] / G.FOCAL, 1.0 / G.FOCAL
elif a.config == "c3":
    p = G.config_c3(a.idx); kind, A, B, me = "fundamental", p["x1"] / G.FOCAL, p["x2"] / G.FOCAL, 1.0 / G.FOCAL
else:
    p = G.config_c4(a.idx); kind, A, B, me = "homography", p["x1"] / G.FOCAL, p["x2"] / G.FOCAL, 1.0 / G.FOCAL
ro = cabi.RansacOpt(**p["ransac"])
for i in range(a.reps):
    t0 = time.perf_counter()
    r = cabi.ransac(kind, A, B, ro, me, rfc=p.get("real_focal_check", False))
    dt = time.perf_counter() - t0
    c = r["counters"]
    print(f"rep {i}: {dt*1e3:.3f} ms  its={r['stats']['iterations']} hyp={c['hypotheses']} launches={c['gpu_launches']} "
          f"k_hyp={c['gpu_seconds']*1e3:.3f} ms lo_wait={c['lo_seconds']*1e3:.3f} ms evaluated={c['samples_evaluated']} "
          f"models_eval={c['models_evaluated']} d2h={c['d2h_bytes']} h2d={c['h2d_bytes']}")
