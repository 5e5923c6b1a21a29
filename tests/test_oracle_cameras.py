"""CPU-only: the oracle's camera models and tangent-Sampson refiner satisfy the property tests the reference holds
for them — tests/camera_models_test.cc:109-320 (project(unproject(x)) == x to 1e-6, analytic projection Jacobian vs
central differences to 1e-6, unprojection Jacobian to 1e-4, on the example cameras of tests/example_cameras.h:31-38)
and tests/optim_relative_test.cc:241-312 (tangent-Sampson Jacobian vs finite differences, refinement reaches a
small gradient)."""
import numpy as np
import plo_py as P
import pytest

# tests/example_cameras.h:31-38 (the six models on the path): (model, width, height, params)
EXAMPLE_CAMERAS = [
    ("SIMPLE_RADIAL", 1936, 1296, [2425.85, 932.383, 628.265, -0.0397695]),
    ("PINHOLE", 6214, 4138, [3425.62, 3426.29, 3118.41, 2069.07]),
    ("SIMPLE_PINHOLE", 6214, 4138, [3425.62, 3118.41, 2069.07]),
    ("RADIAL", 1936, 1296, [2425.85, 932.38, 629.325, -0.04012, 0.00123]),
    ("OPENCV", 3200, 2400, [2575.94, 2608.29, 1599.26, 1257.13, 0.141865, -0.465301, 0, 0]),
    ("OPENCV", 1024, 768, [868.993378, 866.063001, 525.942323, 420.042529, -0.399431, 0.188924, 0.000153, 0.000571]),
]


def _grid(w, h, step=5):
    ij = np.arange(20, 81, step) / 100.0
    g = np.stack(np.meshgrid(ij * w, ij * h, indexing="ij"), -1).reshape(-1, 2)
    return g


def _pp(model, params):
    return np.array(params[2:4] if model in ("PINHOLE", "OPENCV") else params[1:3])


@pytest.mark.parametrize("model,w,h,params", EXAMPLE_CAMERAS)
def test_project_unproject_round_trip(model, w, h, params):
    cam = (model, params)
    pts = np.vstack([_grid(w, h), _pp(model, params) + 1e-6 * np.array([[i, j] for i in (-1, 0, 1) for j in (-1, 0, 1)])])
    d, M = P.camera_unproject_with_jac(cam, pts)
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-12)
    xp, J, xp_plain = P.camera_project_with_jac(cam, d)
    assert np.abs(xp - pts).max() < 1e-6          # camera_models_test.cc:236
    assert np.abs(xp_plain - pts).max() < 1e-6
    # 2D unproject == hnormalized bearing (camera_models.h:98-102)
    u2 = P.camera_unproject2(cam, pts)
    assert np.abs(u2 - d[:, :2] / d[:, 2:3]).max() < 1e-12


@pytest.mark.parametrize("model,w,h,params", EXAMPLE_CAMERAS)
def test_projection_and_unprojection_jacobians(model, w, h, params):
    cam = (model, params)
    pts = _grid(w, h, 10)
    d, M = P.camera_unproject_with_jac(cam, pts)
    _, J, _ = P.camera_project_with_jac(cam, d)
    hstep = 1e-8
    Jfd = np.zeros_like(J)
    for c in range(3):
        e = np.zeros(3)
        e[c] = hstep
        Jfd[:, :, c] = (P.camera_project_with_jac(cam, d + e)[2] - P.camera_project_with_jac(cam, d - e)[2]) / (2 * hstep)
    err = np.linalg.norm((J - Jfd).reshape(len(pts), -1), axis=1) / np.linalg.norm(Jfd.reshape(len(pts), -1), axis=1)
    assert err.max() < 1e-6, err.max()            # camera_models_test.cc:238-242
    Mfd = np.zeros_like(M)
    for c in range(2):
        e = np.zeros(2)
        e[c] = hstep
        Mfd[:, :, c] = (P.camera_unproject_with_jac(cam, pts + e)[0] - P.camera_unproject_with_jac(cam, pts - e)[0]) / (2 * hstep)
    err = np.linalg.norm((M - Mfd).reshape(len(pts), -1), axis=1) / np.linalg.norm(Mfd.reshape(len(pts), -1), axis=1)
    assert err.max() < 1e-4, err.max()            # camera_models_test.cc:268-272


def test_focal_and_null_camera():
    assert P.camera_focal(None) == 1.0
    assert P.camera_focal(("PINHOLE", [100.0, 300.0, 1, 2])) == 200.0
    assert P.camera_focal(("RADIAL", [123.0, 1, 2, 0.1, 0.01])) == 123.0
    x = np.array([[0.3, -0.2], [0.0, 0.0]])
    assert np.array_equal(P.camera_unproject2(None, x), x)   # NULL: homogeneous() then hnormalized() is exact


def _scene(cam, n, seed, noise_px):
    """optim_test_utils.h setup_scene analogue: random pose, 3D points in front of both cameras, projected."""
    rng = np.random.default_rng(seed)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    # small rotation so that points stay in both fields of view
    th = 0.1 * rng.normal(size=3)
    K = np.array([[0, -th[2], th[1]], [th[2], 0, -th[0]], [-th[1], th[0], 0]])
    R = np.eye(3) + K + 0.5 * K @ K
    U, _, Vt = np.linalg.svd(R)
    R = U @ Vt
    t = rng.normal(size=3)
    t /= np.linalg.norm(t)
    X = np.c_[rng.uniform(-0.5, 0.5, n), rng.uniform(-0.4, 0.4, n), rng.uniform(3.0, 8.0, n)]
    X2 = X @ R.T + t
    x1 = P.camera_project_with_jac(cam, X)[2] + rng.normal(0, noise_px, (n, 2)) * (noise_px > 0)
    x2 = P.camera_project_with_jac(cam, X2)[2] + rng.normal(0, noise_px, (n, 2)) * (noise_px > 0)
    tr = np.trace(R)
    qw = np.sqrt(max(0, 1 + tr)) / 2
    qv = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (4 * qw)
    return x1, x2, np.r_[qw, qv, t]


def test_tangent_sampson_zero_at_ground_truth_and_refinement():
    model, w, h, params = EXAMPLE_CAMERAS[3]   # the RADIAL camera of optim_relative_test.cc:243
    cam = (model, params)
    x1, x2, pose = _scene(cam, 25, 3, 0.0)
    d1, M1 = P.camera_unproject_with_jac(cam, x1)
    d2, M2 = P.camera_unproject_with_jac(cam, x2)
    s, cnt = P.score_tangent(pose, d1, d2, M1, M2, 1.0)
    assert cnt == 25 and s < 1e-12
    m, bs = P.refine_relpose_tangent(pose, d1, d2, M1, M2, P.BundleOpt(loss_type="TRIVIAL", max_iterations=5))
    assert bs[1] < 1e-12 and bs[6] < 1e-4   # pixel units: |J| ~ focal

    # noisy refinement in focal-normalised units (optim_relative_test.cc:277-312)
    x1, x2, pose = _scene(cam, 25, 4, 2.0)
    f = params[0]
    cs = (model, [1.0, params[1] / f, params[2] / f, params[3], params[4]])
    d1, M1 = P.camera_unproject_with_jac(cs, x1 / f)
    d2, M2 = P.camera_unproject_with_jac(cs, x2 / f)
    m, bs = P.refine_relpose_tangent(pose, d1, d2, M1, M2, P.BundleOpt(step_tol=1e-12))
    assert bs[2] <= bs[1] and bs[6] < 1e-8, bs


def test_tangent_sampson_jacobian_matches_finite_differences():
    """For a pinhole camera of focal 1 the tangent-Sampson error is the Sampson error of the same constraint scaled by
    1/(|x1~||x2~|): scores agree and both refiners converge to the same pose up to second-order terms."""
    cam = ("PINHOLE", [1.0, 1.0, 0.0, 0.0])
    x1, x2, pose = _scene(cam, 40, 7, 0.0)
    rng = np.random.default_rng(0)
    x1 = x1 + rng.normal(0, 1e-3, x1.shape)
    x2 = x2 + rng.normal(0, 1e-3, x2.shape)
    d1, M1 = P.camera_unproject_with_jac(cam, x1)
    d2, M2 = P.camera_unproject_with_jac(cam, x2)
    st, ct = P.score_tangent(pose, d1, d2, M1, M2, 1e-4)
    ss, cs = P.score("relpose", pose, x1, x2, 1e-4)
    assert ct == cs and abs(st - ss) < 1e-9 * max(1.0, ss)
    mt, bt = P.refine_relpose_tangent(pose, d1, d2, M1, M2, P.BundleOpt(loss_type="TRIVIAL"))
    ms, bs = P.refine("relpose", pose, x1, x2, P.BundleOpt(loss_type="TRIVIAL"))
    assert np.abs(mt - ms).max() < 1e-4 and abs(bt[2] - bs[2]) < 1e-3 * bs[2]


def test_estimate_relative_pose_with_distortion_cameras():
    """robust.cc:242-314 on a distorted camera: both branches recover the pose and agree on the inlier set size."""
    model, w, h, params = EXAMPLE_CAMERAS[3]
    cam = (model, params)
    x1, x2, pose = _scene(cam, 300, 11, 0.5)
    rng = np.random.default_rng(1)
    out = rng.permutation(300)[:90]
    x2[out] = np.c_[rng.uniform(0.2 * w, 0.8 * w, 90), rng.uniform(0.2 * h, 0.8 * h, 90)]
    ro = P.RansacOpt(max_iterations=2000, min_iterations=200, seed=3)
    a = P.estimate("relpose", x1, x2, ro, P.BundleOpt(), 2.0, cam, cam)
    b = P.estimate("relpose", x1, x2, ro, P.BundleOpt(), 2.0, cam, cam, tangent_sampson=True)
    for r in (a, b):
        assert r["stats"]["num_inliers"] >= 190
        q, t = r["model"][:4], r["model"][4:]
        assert min(np.abs(q - pose[:4]).max(), np.abs(q + pose[:4]).max()) < 5e-3
        assert np.abs(t / np.linalg.norm(t) - pose[4:]).max() < 5e-2
    assert abs(a["stats"]["num_inliers"] - b["stats"]["num_inliers"]) <= 6
