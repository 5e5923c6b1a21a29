"""Regenerates tests/golden/reference_sources.json: outputs OF THE REFERENCE'S OWN SOURCES.

PoseLib's sources for the whole path (robust/ransac.cc, estimators, solvers, scorers, bundle, robust.cc) are compiled
unmodified where they lie under /root/reference on top of mini-Eigen (oracle/_ref/libplref2.so, `make -C oracle ref2`,
DESIGN.md §2) and run on the cases of tests/golden/reference_cases.py.  /root/reference exists in the build container
only, so the outputs are committed: tests/test_golden_reference.py holds the oracle (CPU) and the CUDA path (GPU box)
to them.

    python tests/golden/make_reference_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import plo_py as P  # noqa: E402
import reference_cases as RC  # noqa: E402


def main():
    if not P.ref2_available():
        raise SystemExit("needs /root/reference (oracle/_ref/libplref2.so)")
    out = {"_about": "outputs of PoseLib's own sources (a69263d) run on mini-Eigen; regenerate with "
                     "tests/golden/make_reference_golden.py", "cases": {}}
    for name, make in RC.CASES.items():
        case = make()
        with P.reference_sources():
            r = RC.run(P, case)
        out["cases"][name] = {
            "kind": case["kind"], "n": int(len(case["a"])),
            "iterations": r["stats"]["iterations"], "refinements": r["stats"]["refinements"],
            "num_inliers": r["stats"]["num_inliers"], "model_score": float(r["stats"]["model_score"]).hex(),
            "inliers": RC.pack_mask(r["inliers"]),
            "model": [float(v).hex() for v in np.asarray(r["model"], dtype=np.float64).reshape(-1)],
        }
        print(name, r["stats"]["iterations"], r["stats"]["num_inliers"])
    out["refine"] = {}
    for kind, loss in RC.REFINE_CELLS:
        cell = []
        for m0, a, b, kw in RC.refine_cases(kind, loss):
            with P.reference_sources():
                m, bs = P.refine(kind, m0, a, b, P.BundleOpt(**kw))
            cell.append({"iterations": int(bs[0]), "initial_cost": float(bs[1]).hex(), "cost": float(bs[2]).hex(),
                         "model": [float(v).hex() for v in np.asarray(m, dtype=np.float64).reshape(-1)]})
        out["refine"][f"{kind}/{loss}"] = cell
    out["solvers"] = {}
    for name in RC.SOLVERS:
        a, b = RC.solver_instances(name)
        sols = []
        for i in range(len(a)):
            with P.reference_sources():
                r = RC.solve_one(P, name, a[i], b[i])
            sols.append([float(v).hex() for v in r.reshape(-1)])
        out["solvers"][name] = sols
    json.dump(out, open(os.path.join(HERE, "reference_sources.json"), "w"), indent=1)
    print("wrote", len(out["cases"]), "cases and", 3 * len(out["refine"]), "refinements")


if __name__ == "__main__":
    main()
