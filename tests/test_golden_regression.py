"""The committed oracle regression fixtures (tests/golden/): the oracle must still reproduce them exactly (sample
tables, trajectory counters, score bits, inlier masks, model bits).  The GPU parity tests compare the CUDA path with the
live oracle, so a drift of the oracle between rounds would otherwise go unnoticed."""
import importlib.util
import json
import os

import plo_py as P
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(spec)
spec.loader.exec_module(MG)
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_regression.json")))["cases"]


@pytest.mark.parametrize("name,kind,gen,kw,me", MG.CASES, ids=[c[0] for c in MG.CASES])
def test_oracle_reproduces_committed_fixtures(name, kind, gen, kw, me):
    c = GOLD[name]
    a, b, r = MG.run_case(kind, gen(), kw, me)
    k = {"pnp": 3, "relpose": 5, "fundamental": 7, "homography": 4}[kind]
    assert P.sample_table(len(a), k, P.RansacOpt(**kw), 8).tolist() == c["first_samples"]
    for q in ("iterations", "refinements", "num_inliers"):
        assert r["stats"][q] == c["stats"][q], q
    for q in ("samples", "hypotheses", "lo_calls"):
        assert r["counters"][q] == c["counters"][q], q
    assert float(r["stats"]["model_score"]).hex() == c["model_score"]
    assert "".join("1" if v else "0" for v in r["inliers"]) == c["inliers"]
    assert MG.hexlist(r["model"]) == c["model"]
