"""Writes tests/golden/c5_oracle.npz: the CPU oracle's result for every problem of BASELINE config 5 (4096 independent
problems, even index: p3p C1-type, odd index: 5pt C2-type; the same generator calls as bench.py's c5_problems) — iterations,
refinements, inlier count, CRC32 of the inlier mask and the model.  tests/test_zz_c5_full.py holds the CUDA path to it on
the GPU box, problem by problem.  Run on the CPU (≈ 2 minutes on 8 cores):  python tests/golden/make_c5_golden.py"""
import os
import sys
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import plo_py as P  # noqa: E402
from poselib_b200 import problem_generator as G  # noqa: E402

COUNT = 4096


def solve(i):
    F = G.FOCAL
    if i % 2 == 0:
        p = G.abspose_problem(200, 0.5, 5, i)
        r = P.ransac("pnp", p["x"] / F, p["X"], P.RansacOpt(max_iterations=1000, min_iterations=1000), 12.0 / F)
    else:
        p = G.relpose_problem(10000, 0.3, 5, i)
        r = P.ransac("relpose", p["x1"] / F, p["x2"] / F, P.RansacOpt(max_iterations=100000, min_iterations=1000), 1.0 / F)
    st = r["stats"]
    mask = np.asarray(r["inliers"], dtype=np.uint8)
    return (st["iterations"], st["refinements"], st["num_inliers"], zlib.crc32(mask.tobytes()), np.asarray(r["model"], dtype=np.float64))


def main():
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:  # ctypes releases the GIL inside the oracle call
        res = list(ex.map(solve, range(COUNT)))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c5_oracle.npz"),
                        iterations=np.array([r[0] for r in res], dtype=np.int64),
                        refinements=np.array([r[1] for r in res], dtype=np.int64),
                        num_inliers=np.array([r[2] for r in res], dtype=np.int64),
                        mask_crc=np.array([r[3] for r in res], dtype=np.uint32),
                        model=np.stack([r[4] for r in res]))
    print("wrote", COUNT, "results")


if __name__ == "__main__":
    main()
