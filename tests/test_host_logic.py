"""CPU-only: the host side of the engine (it needs no device): the sampler that feeds every round and the
dynamic-iteration formula that sizes the rounds, against the oracle and against the known answers of the reference's
own tests (tests/ransac_test.cc:38-69) / SURVEY.md appendix A."""
import math

import numpy as np
import plo_py as P
import pytest

from poselib_b200 import cabi


@pytest.mark.parametrize("n,k", [(10000, 5), (200, 3), (5000, 7), (20000, 4), (7, 7), (6, 5), (33, 4)])
@pytest.mark.parametrize("seed", [0, 1, 12345, 2**40 + 7])
def test_host_sampler_matches_oracle_uniform(n, k, seed):
    g = cabi.host_sample_table(n, k, cabi.RansacOpt(seed=seed), 3000)
    o = P.sample_table(n, k, P.RansacOpt(seed=seed), 3000)
    assert np.array_equal(g, o)
    assert all(len(set(r)) == k for r in g[:200]) and g.max() < n


@pytest.mark.parametrize("n,k,budget", [(5000, 7, 100000), (400, 5, 300), (50, 4, 40), (64, 3, 100000)])
def test_host_sampler_matches_oracle_prosac(n, k, budget):
    kw = dict(seed=3, progressive_sampling=True, max_prosac_iterations=budget)
    g = cabi.host_sample_table(n, k, cabi.RansacOpt(**kw), 2000)
    o = P.sample_table(n, k, P.RansacOpt(**kw), 2000)
    assert np.array_equal(g, o)
    # while the PROSAC budget lasts the last index is the newest point of the growing subset (sampling.cc:86-102)
    head = g[: min(budget - 1, 2000)]
    assert np.all(np.diff(head[:, -1].astype(np.int64)) >= 0)


def test_first_sample_known_answer():
    # seed 0, N = 10000 -> first 5pt sample {767, 6356, 5535, 6620, 4395}
    g = cabi.host_sample_table(10000, 5, cabi.RansacOpt(seed=0), 1)
    assert g[0].tolist() == [767, 6356, 5535, 6620, 4395]


def test_dynamic_max_iter_known_answers():
    # SURVEY.md appendix A.1 (same formula as tests/ransac_test.cc:52-69): nominal inlier counts of the BASELINE configs
    f = cabi.host_dynamic_max_iter
    assert f(3000, 10000, 5, 0.9999, 3.0, 1000, 100000) == 11384
    assert f(1000, 5000, 7, 0.9999, 3.0, 1000, 100000) == 100000
    assert f(12000, 20000, 4, 0.9999, 3.0, 1000, 100000) == 1000
    assert f(100, 200, 3, 0.9999, 3.0, 1000, 1000) == 1000
    assert f(0, 100, 5, 0.9999, 3.0, 10, 777) == 777
    rng = np.random.default_rng(0)
    for _ in range(300):
        nd = int(rng.integers(8, 30000))
        ni = int(rng.integers(0, nd + 1))
        k = int(rng.choice([3, 4, 5, 7]))
        sp = float(rng.choice([0.99, 0.999, 0.9999]))
        mult = float(rng.choice([1.0, 3.0]))
        mn, mx = int(rng.integers(0, 2000)), int(rng.integers(1, 200000))
        assert f(ni, nd, k, sp, mult, mn, mx) == P.compute_dynamic_max_iter(ni, nd, k, math.log(1.0 - sp), mult, mn, mx)
