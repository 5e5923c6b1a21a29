"""Case table shared by tests/golden/make_reference_golden.py (which runs the REFERENCE'S OWN SOURCES on it), the CPU test
(oracle vs the committed fixtures) and the GPU test (CUDA path vs the committed fixtures).  Every case repeats, input for
input and option for option, a case of tests/test_gpu_parity.py."""
import numpy as np

from poselib_b200 import problem_generator as G

CAMT = (G.FOCAL, G.FOCAL, 0.0, 0.0)


def _pnp(n, ratio, its, seed):
    p = G.abspose_problem(n, ratio, 1, seed)
    return dict(api="ransac", kind="pnp", a=p["x"] / G.FOCAL, b=p["X"], me=12.0 / G.FOCAL,
                kw=dict(max_iterations=its, min_iterations=min(its, 1000), seed=seed))


def _relpose(n, ratio, seed):
    p = G.relpose_problem(n, ratio, 2, seed)
    return dict(api="ransac", kind="relpose", a=p["x1"] / G.FOCAL, b=p["x2"] / G.FOCAL, me=1.0 / G.FOCAL,
                kw=dict(max_iterations=100000, min_iterations=1000, seed=seed))


def _fundamental(prosac, rfc, seed):
    p = G.relpose_problem(2000, 0.3, 3, seed, prosac_sorted=prosac)
    return dict(api="ransac", kind="fundamental", a=p["x1"] / 500.0, b=p["x2"] / 500.0, me=1.0 / 500.0, rfc=rfc,
                kw=dict(max_iterations=20000, min_iterations=1000, seed=seed, progressive_sampling=prosac,
                        max_prosac_iterations=5000))


def _homography(n, ratio, seed):
    p = G.homography_problem(n, ratio, 4, seed)
    return dict(api="ransac", kind="homography", a=p["x1"] / G.FOCAL, b=p["x2"] / G.FOCAL, me=1.0 / G.FOCAL,
                kw=dict(max_iterations=100000, min_iterations=1000, seed=seed))


def _initial_model():
    p = G.relpose_problem(800, 0.5, 2, 11)
    return dict(api="ransac", kind="relpose", a=p["x1"] / G.FOCAL, b=p["x2"] / G.FOCAL, me=1.0 / G.FOCAL,
                init=np.r_[p["q_gt"], p["t_gt"]],
                kw=dict(max_iterations=2000, min_iterations=50, seed=5, score_initial_model=True))


def _est_pnp():
    p = G.config_c1(3)
    return dict(api="estimate", kind="pnp", a=p["x"], b=p["X"], me=p["max_error"], cams=1, kw=dict(p["ransac"]))


def _est_relpose():
    p = G.relpose_problem(3000, 0.4, 2, 5)
    return dict(api="estimate", kind="relpose", a=p["x1"], b=p["x2"], me=1.0, cams=2,
                kw=dict(max_iterations=20000, min_iterations=500, seed=2))


def _est_fundamental():
    p = G.relpose_problem(2000, 0.4, 3, 6, prosac_sorted=True)
    return dict(api="estimate", kind="fundamental", a=p["x1"], b=p["x2"], me=1.0, cams=0, rfc=True,
                kw=dict(max_iterations=20000, min_iterations=500, seed=2, progressive_sampling=True))


def _est_homography():
    p = G.homography_problem(3000, 0.6, 4, 7)
    return dict(api="estimate", kind="homography", a=p["x1"], b=p["x2"], me=1.0, cams=0,
                kw=dict(max_iterations=20000, min_iterations=500, seed=2))


# ---- cameras with distortion / tangent Sampson (rows N3 / N1): the cases of test_estimate_relative_pose_camera_prestep_on_device,
# test_estimate_absolute_pose_distortion_cameras and test_tangent_sampson_ransac_matches_oracle, input for input -------------
DISTORTION_CAMERAS = [  # tests/example_cameras.h:31-38 of the reference, principal point moved to the image centre
    ("SIMPLE_RADIAL", [1100.0, 30.0, -20.0, -0.0397695]),
    ("RADIAL", [1050.0, -15.0, 25.0, -0.04012, 0.00123]),
    ("OPENCV", [1020.0, 990.0, 12.0, -8.0, 0.0141865, -0.0465301, 0.0005, -0.0003]),
    ("OPENCV", [868.993378, 866.063001, 5.9, -4.0, -0.399431, 0.188924, 0.000153, 0.000571]),
    ("SIMPLE_PINHOLE", [950.0, 3.0, 4.0]),
]


def _distort(cam, x_px):
    """pixel observations of the synthetic pinhole camera (f = G.FOCAL, pp = 0) re-imaged by `cam` (oracle projection,
    pinned bit for bit to the reference's camera code)."""
    import plo_py
    X = np.c_[np.asarray(x_px) / G.FOCAL, np.ones(len(x_px))]
    return plo_py.camera_project_with_jac(cam, X)[2]


def _cam_relpose(i, second=None):
    p = G.relpose_problem(2500, 0.45, 31, 2)
    c1 = DISTORTION_CAMERAS[i]
    c2 = DISTORTION_CAMERAS[second] if second is not None else c1
    return dict(api="estimate", kind="relpose", a=_distort(c1, p["x1"]), b=_distort(c2, p["x2"]), me=1.5,
                cam_specs=[c1, c2], kw=dict(max_iterations=20000, min_iterations=300, seed=4))


def _cam_pnp(i):
    p = G.abspose_problem(1500, 0.5, 32, 4)
    c = DISTORTION_CAMERAS[i]
    return dict(api="estimate", kind="pnp", a=_distort(c, p["x"]), b=p["X"], me=4.0, cam_specs=[c],
                kw=dict(max_iterations=5000, min_iterations=300, seed=1))


def _tangent(i, api):
    p = G.relpose_problem(2000, 0.4, 33, 1)
    c = DISTORTION_CAMERAS[i]
    d = dict(api=api, kind="relpose", a=_distort(c, p["x1"]), b=_distort(c, p["x2"]), me=1.5, cam_specs=[c, c],
             kw=dict(max_iterations=20000, min_iterations=300, seed=6))
    if api == "estimate":
        d["tangent_sampson"] = True
    return d


CASES = {}
for _i in range(5):
    CASES[f"estimate_relpose_camera{_i}"] = lambda i=_i: _cam_relpose(i)
    CASES[f"estimate_pnp_camera{_i}"] = lambda i=_i: _cam_pnp(i)
CASES["estimate_relpose_camera0_and_1"] = lambda: _cam_relpose(0, 1)
for _i in range(4):
    CASES[f"ransac_relpose_cameras{_i}"] = lambda i=_i: _tangent(i, "ransac_relpose_cameras")
    CASES[f"estimate_relpose_tangent_camera{_i}"] = lambda i=_i: _tangent(i, "estimate")
for _s in (0, 1, 2):
    CASES[f"ransac_pnp_200_s{_s}"] = lambda s=_s: _pnp(200, 0.5, 1000, s)
    CASES[f"ransac_pnp_1500_s{_s}"] = lambda s=_s: _pnp(1500, 0.35, 3000, s)
    CASES[f"ransac_relpose_1000_s{_s}"] = lambda s=_s: _relpose(1000, 0.5, s)
    CASES[f"ransac_relpose_10000_s{_s}"] = lambda s=_s: _relpose(10000, 0.3, s)
for _s in (0, 1):
    CASES[f"ransac_fundamental_s{_s}"] = lambda s=_s: _fundamental(False, False, s)
    CASES[f"ransac_fundamental_prosac_rfc_s{_s}"] = lambda s=_s: _fundamental(True, True, s)
    CASES[f"ransac_homography_2000_s{_s}"] = lambda s=_s: _homography(2000, 0.6, s)
    CASES[f"ransac_homography_20000_s{_s}"] = lambda s=_s: _homography(20000, 0.6, s)
CASES["ransac_relpose_initial_model"] = _initial_model
CASES["estimate_pnp_c1"] = _est_pnp
CASES["estimate_relpose_3000"] = _est_relpose
CASES["estimate_fundamental_prosac_rfc"] = _est_fundamental
CASES["estimate_homography_3000"] = _est_homography


# ---- LM refiners: the 48 cases of tests/test_gpu_parity.py::test_lm_refiners_match_oracle, input for input ---------------
def _perturbed(kind, p, rng):
    if kind == "pnp":
        q = p["q_gt"] + rng.normal(0, 0.01, 4)
        return np.r_[q / np.linalg.norm(q), p["t_gt"] + rng.normal(0, 0.02, 3)]
    if kind == "relpose":
        q = p["q_gt"] + rng.normal(0, 0.01, 4)
        t = p["t_gt"] + rng.normal(0, 0.02, 3)
        return np.r_[q / np.linalg.norm(q), t / np.linalg.norm(t)]
    if kind == "fundamental":
        t = p["t_gt"]
        E = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) @ p["R_gt"]
        E = E + rng.normal(0, 0.01, (3, 3))
        return E / np.linalg.norm(E)
    H = p["H_gt"] / np.linalg.norm(p["H_gt"])
    return H + rng.normal(0, 0.002, (3, 3))


def refine_cases(kind, loss):
    """The three problems of one (kind, loss) cell: (model0, a, b, BundleOpt kwargs)."""
    rng = np.random.default_rng(7)
    out = []
    for idx in range(3):
        if kind == "pnp":
            p = G.abspose_problem(500, 0.7, 31, idx)
            a, b, scale = p["x"] / G.FOCAL, p["X"], 12.0 / G.FOCAL
        elif kind == "homography":
            p = G.homography_problem(600, 0.7, 34, idx)
            a, b, scale = p["x1"] / G.FOCAL, p["x2"] / G.FOCAL, 2.0 / G.FOCAL
        else:
            p = G.relpose_problem(600, 0.7, 32, idx)
            a, b, scale = p["x1"] / G.FOCAL, p["x2"] / G.FOCAL, 2.0 / G.FOCAL
        out.append((_perturbed(kind, p, rng), a, b, dict(max_iterations=25, loss_type=loss, loss_scale=scale)))
    return out


REFINE_CELLS = [(k, l) for k in ("pnp", "relpose", "fundamental", "homography") for l in ("TRUNCATED", "CAUCHY", "HUBER", "TRIVIAL")]


def check_refine(model, bstats, gold, model_tol, cost0_rtol, cost_rtol):
    """bstats = [iterations, initial_cost, cost, ...]; F / H up to sign (same comparison as the GPU parity test)."""
    gm = np.array([float.fromhex(v) for v in gold["model"]]).reshape(np.asarray(model).shape)
    c0, c1 = float.fromhex(gold["initial_cost"]), float.fromhex(gold["cost"])
    assert abs(bstats[1] - c0) <= cost0_rtol * abs(c0), (bstats[1], c0)
    assert abs(bstats[2] - c1) <= cost_rtol * abs(c1), (bstats[2], c1)
    m = np.asarray(model, dtype=np.float64)
    err = min(np.abs(m - gm).max(), np.abs(m + gm).max()) if m.ndim == 2 else np.abs(m - gm).max()
    assert err <= model_tol * max(1.0, np.abs(gm).max()), (err, m, gm)
    return err


# ---- minimal solvers: subsets of the instances of test_p3p / test_relpose_7pt / test_homography_4pt_matches_oracle --------
def _noisy_samples(npts, count, seed, outliers=True):
    """bearing samples as the estimators build them from noisy / outlier data (same generator as tests/test_gpu_parity.py)."""
    rng = np.random.default_rng(seed)
    x1s, x2s = [], []
    for i in range(count):
        p = G.relpose_problem(64, 0.5 if outliers else 1.0, config_id=20, problem_idx=seed * 1000 + i)
        idx = rng.choice(64, npts, replace=False)
        a = np.c_[p["x1"][idx] / G.FOCAL, np.ones(npts)]
        b = np.c_[p["x2"][idx] / G.FOCAL, np.ones(npts)]
        x1s.append(a / np.linalg.norm(a, axis=1, keepdims=True))
        x2s.append(b / np.linalg.norm(b, axis=1, keepdims=True))
    return np.array(x1s), np.array(x2s)


def solver_instances(name, n_min=40, n_noisy=12):
    """(a, b) arrays [count, k, 3]: the first n_min minimal instances and the first n_noisy noisy samples of the GPU test."""
    if name in ("p3p", "p3p_lambdatwist"):
        xs, Xs = [], []
        for i in range(n_min):
            x, X, _, _ = G.minimal_abspose(i)
            xs.append(x)
            Xs.append(X)
        for i in range(n_noisy):
            p = G.abspose_problem(50, 0.5, config_id=21, problem_idx=i)
            a = np.c_[p["x"][:3] / G.FOCAL, np.ones(3)]
            xs.append(a / np.linalg.norm(a, axis=1, keepdims=True))
            Xs.append(p["X"][:3])
        return np.array(xs), np.array(Xs)
    k, seed = {"relpose_7pt": (7, 2), "homography_4pt": (4, 3)}[name]
    x1s, x2s = [], []
    for i in range(n_min):
        if name == "relpose_7pt":
            x1, x2, _, _ = G.minimal_relpose(i, 7)
        else:
            x1, x2, _ = G.minimal_homography(i)
        x1s.append(x1)
        x2s.append(x2)
    a, b = _noisy_samples(k, 300, seed)  # the GPU test draws 300; the stream must be consumed identically
    return np.concatenate([np.array(x1s), a[:n_noisy]]), np.concatenate([np.array(x2s), b[:n_noisy]])


SOLVERS = ("p3p", "p3p_lambdatwist", "relpose_7pt", "homography_4pt")


def solve_one(api, name, a, b):
    """One instance through the oracle-style wrapper; -> array of solutions (k, 7) / (k, 3, 3)."""
    if name == "p3p":
        return np.asarray(api.p3p(a, b))
    if name == "p3p_lambdatwist":
        return np.asarray(api.p3p_lambdatwist(a, b))
    if name == "relpose_7pt":
        return np.asarray(api.relpose_7pt(a, b))
    n, H = api.homography_4pt(a, b)
    return np.asarray(H)[None] if n else np.zeros((0, 3, 3))


def run(api, case):
    """api: oracle/plo_py (also inside `with plo_py.reference_sources()`) or poselib_b200.cabi — same call surface."""
    ro = api.RansacOpt(**case["kw"])
    extra = {}
    if "rfc" in case:
        extra["rfc"] = case["rfc"]
    if "init" in case:
        extra["init"] = case["init"]
    if case["api"] == "ransac":
        return api.ransac(case["kind"], case["a"], case["b"], ro, case["me"], **extra)
    if "cam_specs" in case:  # (model name, params): the oracle wrappers take the tuple, the C-ABI binding a Camera struct
        cams = [api.Camera(*c) if hasattr(api, "Camera") else c for c in case["cam_specs"]]
        if case["api"] == "ransac_relpose_cameras":
            return api.ransac_relpose_cameras(case["a"], case["b"], cams[0], cams[1], ro, case["me"])
        if case.get("tangent_sampson"):
            extra["tangent_sampson"] = True
        return api.estimate(case["kind"], case["a"], case["b"], ro, api.BundleOpt(), case["me"], *cams, **extra)
    cam = api.Camera("PINHOLE", CAMT) if hasattr(api, "Camera") else CAMT
    cams = [cam] * case["cams"]
    return api.estimate(case["kind"], case["a"], case["b"], ro, api.BundleOpt(), case["me"], *cams, **extra)


def pack_mask(m):
    return np.packbits(np.asarray(m, dtype=np.uint8) != 0).tobytes().hex()


def check(result, gold, kind, model_tol, score_rtol):
    """Discrete outputs exact; score and model within the stated tolerances (F / H up to sign, |t| of a relative pose
    up to its gauge — the same normalisations as tests/test_gpu_parity.py::_same_trajectory)."""
    for k in ("iterations", "refinements", "num_inliers"):
        assert result["stats"][k] == gold[k], (k, result["stats"][k], gold[k])
    assert pack_mask(result["inliers"]) == gold["inliers"]
    gs = float.fromhex(gold["model_score"])
    assert abs(result["stats"]["model_score"] - gs) <= score_rtol * abs(gs), (result["stats"]["model_score"], gs)
    gm = np.array([float.fromhex(v) for v in gold["model"]]).reshape(np.asarray(result["model"]).shape)
    rm = np.asarray(result["model"], dtype=np.float64)
    if rm.ndim == 2:
        err = min(np.abs(rm - gm).max(), np.abs(rm + gm).max())
    else:
        if kind == "relpose":
            rm = np.r_[rm[:4], rm[4:] / max(np.linalg.norm(rm[4:]), 1e-300)]
            gm = np.r_[gm[:4], gm[4:] / max(np.linalg.norm(gm[4:]), 1e-300)]
        err = np.abs(rm - gm).max()
    assert err <= model_tol * np.abs(gm).max(), (err, rm, gm)
    return err / np.abs(gm).max()
