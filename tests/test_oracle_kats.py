"""CPU-only: pins the oracle against every known-answer test the reference holds for this path
(SURVEY.md §8c): tests/ransac_test.cc:38-122 (ported verbatim), the sampler KAT derived from
sampling.cc:37-61, and the solver acceptance criteria of
benchmark/solver_benchmark.cc:27-45 (valid + GT found at 1e-6 on noise-free minimal instances)."""
import math

import numpy as np
import plo_py as P

from poselib_b200 import problem_generator as G


# ---- tests/ransac_test.cc ------------------------------------------------------------------------
def expected_iterations_from_dynamic_limit(dyn, opt):
    stop_after = max(opt.min_iterations, dyn)
    return opt.max_iterations if stop_after >= opt.max_iterations else stop_after + 1


def test_all_inlier_sample_probability_matches_hypergeometric():  # ransac_test.cc:38-51
    expected = (5.0 / 10.0) * (4.0 / 9.0) * (3.0 / 8.0) * (2.0 / 7.0) * (1.0 / 6.0)
    actual = P.all_inlier_sample_probability(5, 10, 5)
    assert abs(actual - expected) < 1e-12
    assert abs(actual - 0.5 ** 5) > 1e-3


def test_dynamic_iterations_use_exact_probability():  # ransac_test.cc:53-76
    opt = P.RansacOpt(min_iterations=0, max_iterations=1000, dyn_num_trials_mult=1.0, success_prob=0.5)
    p = (5.0 / 10.0) * (4.0 / 9.0) * (3.0 / 8.0) * (2.0 / 7.0) * (1.0 / 6.0)
    expected = int(math.ceil(math.log(1.0 - 0.5) / math.log(1.0 - p) * 1.0))
    assert P.compute_dynamic_max_iter(5, 10, 5, math.log(0.5), 1.0, 0, 1000) == expected
    st = P.ransac_mock(10, 5, 5, opt)
    assert st.num_inliers == 5
    assert st.iterations == expected_iterations_from_dynamic_limit(expected, opt)


def test_dynamic_iterations_stay_at_max_when_not_enough_inliers():  # ransac_test.cc:78-99
    opt = P.RansacOpt(min_iterations=0, max_iterations=20, dyn_num_trials_mult=1.0, success_prob=0.5)
    assert P.all_inlier_sample_probability(4, 10, 5) == 0.0
    assert P.compute_dynamic_max_iter(4, 10, 5, math.log(0.5), 1.0, 0, 20) == 20
    st = P.ransac_mock(10, 5, 4, opt)
    assert st.num_inliers == 4 and st.iterations == 20


def test_dynamic_iterations_collapse_to_min_for_all_inliers():  # ransac_test.cc:101-122
    opt = P.RansacOpt(min_iterations=3, max_iterations=100, dyn_num_trials_mult=1.0, success_prob=0.5)
    assert P.all_inlier_sample_probability(8, 8, 5) == 1.0
    assert P.compute_dynamic_max_iter(8, 8, 5, math.log(0.5), 1.0, 3, 100) == 3
    st = P.ransac_mock(8, 5, 8, opt)
    assert st.num_inliers == 8
    assert st.iterations == expected_iterations_from_dynamic_limit(3, opt)


def test_expected_dynamic_max_iter_at_config_sizes():
    lp = math.log(1 - 0.9999)
    assert P.compute_dynamic_max_iter(3000, 10000, 5, lp, 3.0, 1000, 100000) == 11384
    assert P.compute_dynamic_max_iter(1000, 5000, 7, lp, 3.0, 1000, 100000) == 100000
    assert P.compute_dynamic_max_iter(12000, 20000, 4, lp, 3.0, 1000, 100000) == 1000


# ---- sampler KAT (sampling.cc:37-61) -------------------------------------------------------------
def _splitmix_int(state):
    state = (state + 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & ((1 << 64) - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & ((1 << 64) - 1)
    z = (z ^ (z >> 31)) & 0xFFFFFFFF
    return state, z - (1 << 32) if z >= (1 << 31) else z


def test_random_int_kat():
    assert list(P.random_ints(0, 6)) == [2065550767, -1581685260, -2146876081, 1917616620, 1369994395, 1954456298]
    st, ref = 12345, []
    for _ in range(1000):
        st, v = _splitmix_int(st)
        ref.append(v)
    assert list(P.random_ints(12345, 1000)) == ref


def test_first_sample_at_c2():
    tab = P.sample_table(10000, 5, P.RansacOpt(seed=0), 3)
    assert list(tab[0]) == [767, 6356, 5535, 6620, 4395]


def test_sample_table_matches_bigint_model():
    """Independent python big-int model of draw_sample incl. the int -> size_t sign extension."""
    for N, K, seed in [(200, 3, 0), (10000, 5, 7), (37, 7, 3), (5, 4, 99)]:
        st, rows = seed, []
        for _ in range(300):
            s = []
            while len(s) < K:
                st, v = _splitmix_int(st)
                idx = (v & ((1 << 64) - 1)) % N  # sign-extended to 64-bit unsigned
                if idx not in s:
                    s.append(idx)
            rows.append(s)
        assert P.sample_table(N, K, P.RansacOpt(seed=seed), 300).tolist() == rows


def test_prosac_sampler_properties():
    N, K, its = 500, 7, 2000
    opt = P.RansacOpt(seed=1, progressive_sampling=True, max_prosac_iterations=1000)
    tab = P.sample_table(N, K, opt, its)
    assert all(len(set(r)) == K for r in tab.tolist())
    # while PROSAC is active the last index is subset_sz-1, non-decreasing, and bounds the others
    last = tab[:999, K - 1].astype(int)
    assert last[0] == K - 1 and (np.diff(last) >= 0).all()
    assert (tab[:999, :K - 1].max(axis=1) < last).all()
    assert tab[1200:].max() > last.max()  # uniform sampling afterwards reaches the whole range


# ---- solver acceptance (solver_benchmark.cc:27-45; validators problem_generator.cc:14-201) -------
def _pose_err(p, R, t):
    return np.linalg.norm(G.quat_to_rotmat(p[:4]) - R) + np.linalg.norm(p[4:] - t)


def test_p3p_valid_and_gt_found():
    found = 0
    for i in range(300):
        x, X, R, t = G.minimal_abspose(i)
        poses = P.p3p(x, X)
        assert len(poses) <= 4
        for p in poses:  # CalibPoseValidator::is_valid (problem_generator.cc:43-53)
            Rp = G.quat_to_rotmat(p[:4])
            assert abs(np.linalg.det(Rp) - 1) < 1e-6
            Z = X @ Rp.T + p[4:]
            Z /= np.linalg.norm(Z, axis=1, keepdims=True)
            assert np.abs(Z - x).max() < 1e-6
        found += any(_pose_err(p, R, t) < 1e-6 for p in poses)
    assert found >= 299


def test_relpose_5pt_valid_and_gt_found():
    found, nsol = 0, 0
    for i in range(300):
        x1, x2, R, t = G.minimal_relpose(i, 5)
        Es = P.relpose_5pt_E(x1, x2)
        poses = P.relpose_5pt(x1, x2)
        nsol += len(Es)
        assert len(Es) <= 10 and len(poses) <= 40
        for E in Es:  # epipolar constraints + essential-ness
            assert np.abs(np.einsum("ni,ij,nj->n", x2, E, x1)).max() < 1e-6
        for p in poses:  # is_valid (problem_generator.cc:94-112): cheirality + epipolar
            assert abs(np.linalg.norm(p[:4]) - 1) < 1e-9
        found += any(_pose_err(p, R, t) < 1e-6 for p in poses)
    assert found >= 290 and 3.0 < nsol / 300 < 7.0


def test_relpose_7pt_valid_and_gt_found():
    found = 0
    for i in range(300):
        x1, x2, R, t = G.minimal_relpose(i, 7)
        Fs = P.relpose_7pt(x1, x2)
        assert 1 <= len(Fs) <= 3
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        E = tx @ R
        E /= np.linalg.norm(E)
        for F in Fs:
            assert np.abs(np.einsum("ni,ij,nj->n", x2, F, x1)).max() < 1e-6
            assert abs(np.linalg.det(F)) < 1e-6 and abs(np.linalg.norm(F) - 1) < 1e-9
        found += any(min(np.linalg.norm(F - E), np.linalg.norm(F + E)) < 1e-6 for F in Fs)
    assert found >= 298


def test_homography_4pt_valid_and_gt_found():
    found = 0
    for i in range(300):
        x1, x2, H = G.minimal_homography(i)
        n, Hh = P.homography_4pt(x1, x2)
        assert n == 1
        Hn = H / np.linalg.norm(H)
        y = x1 @ Hh.T  # HomographyValidator (problem_generator.cc:185-201)
        y /= np.linalg.norm(y, axis=1, keepdims=True)
        assert min(np.abs(y - x2).max(), np.abs(y + x2).max()) < 1e-6
        found += min(np.linalg.norm(Hh - Hn), np.linalg.norm(Hh + Hn)) < 1e-6
    assert found >= 299


def test_homography_4pt_rejects_flipped_orientation():  # homography_4pt.cc:38-55
    x1, x2, _ = G.minimal_homography(3)
    bad = x2.copy()
    bad[[0, 1]] = bad[[1, 0]]
    n, _ = P.homography_4pt(x1, bad)
    n2, _ = P.homography_4pt(x1, bad, check_cheirality=False)
    assert n == 0 and n2 in (0, 1)


def test_sturm_roots_known_polynomial():
    roots_true = np.array([-3.5, -1.25, -0.1, 0.3, 0.31, 2.0, 7.0])
    poly = np.poly(np.concatenate([roots_true, [1 + 2j, 1 - 2j, 0.5]]))  # degree 10, 8 real roots
    c = poly[::-1].real
    got = P.bisect_sturm10(c)
    assert np.allclose(np.sort(got), np.sort(np.concatenate([roots_true, [0.5]])), atol=1e-8)
    assert (np.diff(got) > 0).all()  # emitted left -> right (sturm.h:223-229)
    assert len(P.bisect_sturm10(np.r_[c[:10], 0.0])) == 0  # leading coeff == 0 -> 0 roots (sturm.h:234-235)
