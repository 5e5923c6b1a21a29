"""CPU-only: the oracle's LM refiners satisfy the property tests the reference holds for them
(tests/optim_{absolute,relative,fundamental,homography}_test.cc): zero residual / gradient at the ground truth
(optim_relative_test.cc:64-91) and "noisy refinement reduces the cost and reaches a small gradient"
(optim_relative_test.cc:124-149, optim_homography_test.cc:168-193, optim_test_utils.h:197-208)."""
import numpy as np
import plo_py as P
import pytest

from poselib_b200 import problem_generator as G


def _scene(kind, idx, noise):
    rng = np.random.default_rng(100 + idx)
    if kind == "pnp":
        p = G.abspose_problem(200, 1.0, 51, idx)
        x = p["x"] / G.FOCAL
        if not noise:
            Z = p["X"] @ p["R_gt"].T + p["t_gt"]
            x = Z[:, :2] / Z[:, 2:3]
        return x, p["X"], np.r_[p["q_gt"], p["t_gt"]]
    if kind == "homography":
        p = G.homography_problem(200, 1.0, 54, idx)
        x1, x2 = p["x1"] / G.FOCAL, p["x2"] / G.FOCAL
        H = p["H_gt"] / p["H_gt"][2, 2]
        if not noise:
            y = np.c_[x1, np.ones(len(x1))] @ H.T
            x2 = y[:, :2] / y[:, 2:3]
        return x1, x2, H
    p = G.relpose_problem(200, 1.0, 52, idx)
    x1, x2 = p["x1"] / G.FOCAL, p["x2"] / G.FOCAL
    R, t = p["R_gt"], p["t_gt"]
    E = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) @ R
    if not noise:  # exact correspondences: project x1 onto the epipolar line in image 2
        l = np.c_[x1, np.ones(len(x1))] @ E.T
        d = (np.sum(np.c_[x2, np.ones(len(x2))] * l, axis=1)) / (l[:, 0] ** 2 + l[:, 1] ** 2)
        x2 = x2 - d[:, None] * l[:, :2]
    if kind == "relpose":
        return x1, x2, np.r_[p["q_gt"], t]
    return x1, x2, E / np.linalg.norm(E)


@pytest.mark.parametrize("kind", ["pnp", "relpose", "fundamental", "homography"])
def test_zero_residual_and_gradient_at_ground_truth(kind):
    for idx in range(3):
        a, b, gt = _scene(kind, idx, noise=False)
        m, bs = P.refine(kind, gt, a, b, P.BundleOpt(max_iterations=10, loss_type="TRIVIAL"))
        assert bs[1] < 1e-18 and bs[2] < 1e-18, bs       # initial / final cost
        assert bs[6] < 1e-9, bs                          # gradient norm at the ground truth
        if np.ndim(gt) == 2:  # projective entities: compare up to scale and sign
            mn, gn = m / np.linalg.norm(m), gt / np.linalg.norm(gt)
            err = min(np.abs(mn - gn).max(), np.abs(mn + gn).max())
        else:
            err = np.abs(m - gt).max()
        assert err < 1e-7


@pytest.mark.parametrize("loss", ["TRIVIAL", "CAUCHY", "TRUNCATED", "HUBER"])
@pytest.mark.parametrize("kind", ["pnp", "relpose", "fundamental", "homography"])
def test_noisy_refinement_reduces_cost(kind, loss):
    rng = np.random.default_rng(5)
    for idx in range(2):
        a, b, gt = _scene(kind, idx, noise=True)
        start = gt + rng.normal(0, 0.005, np.shape(gt))
        if kind in ("pnp", "relpose"):
            start[:4] /= np.linalg.norm(start[:4])
        m, bs = P.refine(kind, start, a, b, P.BundleOpt(max_iterations=100, loss_type=loss, loss_scale=3.0 / G.FOCAL))
        assert bs[2] <= bs[1] and np.isfinite(bs[2]), bs
        if loss == "TRIVIAL" and kind != "fundamental":
            assert bs[6] < 1e-6, bs
