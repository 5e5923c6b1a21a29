"""The `poselib`-compatible Python surface: option-dict handling on CPU, results on the GPU."""
import numpy as np
import pytest

import plo_py as P
from poselib_b200 import problem_generator as G


def test_option_dicts_follow_the_pybind_helpers():
    from poselib_b200 import cabi, pyapi
    kw = pyapi._ransac({"max_iterations": 5, "seed": 7, "progressive_sampling": True, "score_initial_model": True,
                        "success_prob": 0.5})
    assert kw == {"max_iterations": 5, "seed": 7, "progressive_sampling": True, "success_prob": 0.5}  # helpers.h:39
    b = pyapi._bundle({"loss_type": "cauchy", "loss_scale": 2, "max_iterations": 3})
    assert b == {"loss_type": "CAUCHY", "loss_scale": 2.0, "max_iterations": 3}
    with pytest.raises(cabi.PoseLibB200Error):
        pyapi._bundle({"loss_type": "TRUNCATED_LE_ZACH"})
    with pytest.raises(cabi.PoseLibB200Error):
        pyapi._bundle({"damping": "MARQUARDT"})
    with pytest.raises(cabi.PoseLibB200Error):
        pyapi._camera({"model": "OPENCV_FISHEYE", "params": [1, 1, 0, 0, 0, 0, 0, 0]})
    c = pyapi._camera({"model": "SIMPLE_PINHOLE", "width": 640, "height": 480, "params": [500.0, 320.0, 240.0]})
    assert (c.model_id, c.width, c.height, list(c.params)[:3]) == (0, 640, 480, [500.0, 320.0, 240.0])
    c = pyapi._camera({"model": "OPENCV", "params": [900.0, 901.0, 3.0, 4.0, 0.1, -0.2, 1e-3, 2e-3]})
    assert c.model_id == 4 and list(c.params) == [900.0, 901.0, 3.0, 4.0, 0.1, -0.2, 1e-3, 2e-3]


def test_option_factories_return_the_pybind_defaults():
    from poselib_b200 import cabi, pyapi
    r = pyapi.RansacOptions()
    assert r == {"max_iterations": 100000, "min_iterations": 1000, "dyn_num_trials_mult": 3.0, "success_prob": 0.9999,
                 "seed": 0, "progressive_sampling": False, "max_prosac_iterations": 100000}  # types.h:39-50
    assert pyapi.RansacOptions({"seed": 9, "progressive_sampling": True})["seed"] == 9
    b = pyapi.BundleOptions({"loss_type": "huber", "loss_scale": 0.5})
    assert (b["loss_type"], b["loss_scale"], b["max_iterations"], b["lambda_update"], b["damping"]) == ("HUBER", 0.5, 100, "NIELSEN", "LEVENBERG")
    assert sorted(b) == sorted(["max_iterations", "loss_scale", "loss_type", "gradient_tol", "step_tol", "relative_cost_tol",
                                "initial_lambda", "min_lambda", "max_lambda", "lambda_factor", "verbose", "lambda_update",
                                "damping"])  # helpers.h:185-235
    with pytest.raises(cabi.PoseLibB200Error):
        pyapi.BundleOptions({"damping": "MARQUARDT"})
    # the defaults are the ones the C-ABI structs carry
    ro, bo = cabi.RansacOpt(), cabi.BundleOpt()
    for k in ("max_iterations", "min_iterations", "dyn_num_trials_mult", "success_prob", "seed", "max_prosac_iterations"):
        assert getattr(ro, k) == r[k], k
    for k in ("max_iterations", "loss_scale", "gradient_tol", "step_tol", "relative_cost_tol", "initial_lambda", "min_lambda",
              "max_lambda"):
        assert getattr(bo, k) == pyapi.BundleOptions()[k], k


@pytest.mark.gpu
def test_pyapi_matches_oracle_through_the_poselib_call_surface():
    from poselib_b200 import pyapi as poselib
    cam = {"model": "PINHOLE", "width": 2000, "height": 2000, "params": [G.FOCAL, G.FOCAL, 0.0, 0.0]}
    camt = (G.FOCAL, G.FOCAL, 0.0, 0.0)
    p = G.relpose_problem(2000, 0.4, 2, 21)
    opt = {"max_error": 1.0, "ransac": {"max_iterations": 20000, "min_iterations": 500, "seed": 3}}
    pose, info = poselib.estimate_relative_pose(p["x1"], p["x2"], cam, cam, opt)
    o = P.estimate("relpose", p["x1"], p["x2"], P.RansacOpt(max_iterations=20000, min_iterations=500, seed=3),
                   P.BundleOpt(), 1.0, camt, camt)
    assert info["iterations"] == o["stats"]["iterations"] and info["num_inliers"] == o["stats"]["num_inliers"]
    assert info["inliers"] == [bool(v) for v in o["inliers"]]
    assert np.allclose(pose.q, o["model"][:4], atol=1e-7)
    assert np.allclose(pose.R @ pose.R.T, np.eye(3), atol=1e-12) and pose.Rt.shape == (3, 4)
    # initial_pose switches score_initial_model on (relative_pose.cc:25-28)
    pose2, info2 = poselib.estimate_relative_pose(p["x1"], p["x2"], cam, cam, opt, initial_pose=pose)
    assert info2["num_inliers"] >= info["num_inliers"] - 5
    q = G.config_c1(2)
    img, info = poselib.estimate_absolute_pose(q["x"], q["X"], cam, {"max_error": 12.0, "ransac": q["ransac"]})
    o = P.estimate("pnp", q["x"], q["X"], P.RansacOpt(**q["ransac"]), P.BundleOpt(), 12.0, camt)
    assert info["num_inliers"] == o["stats"]["num_inliers"] and np.allclose(img.pose.t, o["model"][4:], atol=1e-6)
    h = G.homography_problem(1500, 0.6, 4, 9)
    H, info = poselib.estimate_homography(h["x1"], h["x2"], {"ransac": {"max_iterations": 5000, "seed": 1}})
    o = P.estimate("homography", h["x1"], h["x2"], P.RansacOpt(max_iterations=5000, seed=1), P.BundleOpt(), 1.0)
    assert info["num_inliers"] == o["stats"]["num_inliers"]
    assert min(np.abs(H - o["model"]).max(), np.abs(H + o["model"]).max()) < 1e-6
    F, info = poselib.estimate_fundamental(p["x1"], p["x2"], {"ransac": {"max_iterations": 5000, "seed": 1}})
    o = P.estimate("fundamental", p["x1"], p["x2"], P.RansacOpt(max_iterations=5000, seed=1), P.BundleOpt(), 1.0)
    assert info["num_inliers"] == o["stats"]["num_inliers"]
    # distorted cameras + RelativePoseOptions.tangent_sampson through the dict interface (relative_pose.cc:15-52)
    camd = {"model": "RADIAL", "width": 2000, "height": 2000, "params": [1050.0, -15.0, 25.0, -0.04012, 0.00123]}
    camo = ("RADIAL", camd["params"])
    X1 = np.c_[p["x1"] / G.FOCAL, np.ones(len(p["x1"]))]
    X2 = np.c_[p["x2"] / G.FOCAL, np.ones(len(p["x2"]))]
    d1, d2 = P.camera_project_with_jac(camo, X1)[2], P.camera_project_with_jac(camo, X2)[2]
    for ts in (False, True):
        optd = {"max_error": 1.5, "tangent_sampson": ts, "ransac": {"max_iterations": 20000, "min_iterations": 500, "seed": 3}}
        pose, info = poselib.estimate_relative_pose(d1, d2, camd, camd, optd)
        o = P.estimate("relpose", d1, d2, P.RansacOpt(max_iterations=20000, min_iterations=500, seed=3), P.BundleOpt(),
                       1.5, camo, camo, tangent_sampson=ts)
        assert info["iterations"] == o["stats"]["iterations"] and info["num_inliers"] == o["stats"]["num_inliers"]
        assert info["inliers"] == [bool(v) for v in o["inliers"]]
        assert np.allclose(pose.q, o["model"][:4], atol=1e-7)
    # solvers
    x, X, R, t = G.minimal_abspose(1)
    assert len(poselib.p3p(x, X)) == len(P.p3p(x, X))
    assert len(poselib.p3p_lambdatwist(x, X)) == len(P.p3p_lambdatwist(x, X))
    x1, x2, R, t = G.minimal_relpose(1, 5)
    assert len(poselib.relpose_5pt(x1, x2)) == len(P.relpose_5pt(x1, x2))
    assert np.allclose(np.array(poselib.essential_matrix_5pt(x1, x2)), P.relpose_5pt_E(x1, x2), atol=1e-12)
    x1, x2, R, t = G.minimal_relpose(1, 7)
    assert np.allclose(np.array(poselib.relpose_7pt(x1, x2)), P.relpose_7pt(x1, x2), atol=1e-9)
    x1, x2, Hgt = G.minimal_homography(1)
    assert len(poselib.homography_4pt(x1, x2)) == 1
