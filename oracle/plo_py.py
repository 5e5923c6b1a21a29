"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_build/libplo.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module.  The product
package (poselib_b200/) never does.  PARITY PARTLY PINNED: the oracle is a restatement of PoseLib's CPU
path without Eigen, not a build of PoseLib; the Eigen-free parts of the reference (sampler, loop templates, univariate
solvers, Sturm) are built into oracle/_ref and pin the corresponding oracle functions bit for bit (tests/test_ref_pins.py),
the reference's own sources for the whole path run on mini-Eigen (oracle/_ref/libplref2.so, tests/test_ref_sources.py)
and pin the transcription of PoseLib's logic end to end; Eigen's own arithmetic (reduction order, decompositions) is
unpinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libplo.so")


class RansacOpt(C.Structure):
    _fields_ = [("max_iterations", C.c_uint64), ("min_iterations", C.c_uint64),
                ("dyn_num_trials_mult", C.c_double), ("success_prob", C.c_double),
                ("seed", C.c_uint64), ("progressive_sampling", C.c_int32),
                ("score_initial_model", C.c_int32), ("max_prosac_iterations", C.c_uint64)]

    def __init__(self, max_iterations=100000, min_iterations=1000, dyn_num_trials_mult=3.0,
                 success_prob=0.9999, seed=0, progressive_sampling=False, score_initial_model=False,
                 max_prosac_iterations=100000):
        super().__init__(max_iterations, min_iterations, dyn_num_trials_mult, success_prob, seed,
                         int(progressive_sampling), int(score_initial_model), max_prosac_iterations)


class RansacStats(C.Structure):
    _fields_ = [("refinements", C.c_uint64), ("iterations", C.c_uint64), ("num_inliers", C.c_uint64),
                ("inlier_ratio", C.c_double), ("model_score", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


LOSS = {"TRIVIAL": 0, "TRUNCATED": 1, "HUBER": 2, "CAUCHY": 3}


class BundleOpt(C.Structure):
    _fields_ = [("max_iterations", C.c_uint64), ("loss_type", C.c_int32), ("pad", C.c_int32),
                ("loss_scale", C.c_double), ("gradient_tol", C.c_double), ("step_tol", C.c_double),
                ("relative_cost_tol", C.c_double), ("initial_lambda", C.c_double),
                ("min_lambda", C.c_double), ("max_lambda", C.c_double)]

    def __init__(self, max_iterations=100, loss_type="CAUCHY", loss_scale=1.0, gradient_tol=1e-12,
                 step_tol=1e-8, relative_cost_tol=1e-10, initial_lambda=1e-3, min_lambda=1e-10,
                 max_lambda=1e10):
        lt = LOSS[loss_type] if isinstance(loss_type, str) else int(loss_type)
        super().__init__(max_iterations, lt, 0, loss_scale, gradient_tol, step_tol, relative_cost_tol,
                         initial_lambda, min_lambda, max_lambda)


class Counters(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("hypotheses", C.c_uint64), ("scored_corrs", C.c_uint64),
                ("lo_calls", C.c_uint64), ("lo_seconds", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only; seconds)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


_DOUBLE_RETURNS = ("plo_score_pnp", "plo_score_relpose", "plo_score_fundamental", "plo_score_homography",
                   "plo_score_tangent", "plo_camera_focal")


def lib():
    global _lib
    if _use_ref2:
        return _Ref2Proxy()
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.plo_all_inlier_sample_probability.restype = C.c_double
        _lib.plo_compute_dynamic_max_iter.restype = C.c_uint64
        for n in ("plo_score_pnp", "plo_score_relpose", "plo_score_fundamental", "plo_score_homography",
                  "plo_ransac_relpose_batch_mt"):
            getattr(_lib, n).restype = C.c_double
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _mask(n):
    m = np.zeros(n, dtype=np.int8)
    return m, m.ctypes.data_as(C.c_char_p)


# ---- sampler ------------------------------------------------------------------------------------
def random_ints(seed, n):
    out = np.zeros(n, dtype=np.int32)
    lib().plo_random_ints(C.c_uint64(seed), n, out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


def sample_table(N, K, opt, iters):
    out = np.zeros((iters, K), dtype=np.uint32)
    lib().plo_sample_table(C.c_uint64(N), C.c_uint64(K), C.byref(opt), C.c_uint64(iters),
                           out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


# ---- loop KATs ----------------------------------------------------------------------------------
def all_inlier_sample_probability(ni, nd, k):
    return lib().plo_all_inlier_sample_probability(C.c_uint64(ni), C.c_uint64(nd), C.c_uint64(k))


def compute_dynamic_max_iter(ni, nd, k, logp, mult, mn, mx):
    return lib().plo_compute_dynamic_max_iter(C.c_uint64(ni), C.c_uint64(nd), C.c_uint64(k), C.c_double(logp),
                                              C.c_double(mult), C.c_uint64(mn), C.c_uint64(mx))


def ransac_mock(nd, k, inl, opt):
    st = RansacStats()
    lib().plo_ransac_mock(C.c_uint64(nd), C.c_uint64(k), C.c_uint64(inl), C.byref(opt), C.byref(st))
    return st


# ---- solvers (unit bearings, row-per-point arrays) ----------------------------------------------
def p3p(x, X):
    xa, xp = _d(x)
    Xa, Xp = _d(X)
    out = np.zeros((4, 7))
    n = lib().plo_p3p(xp, Xp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out[:n]


def p3p_lambdatwist(x, X):
    xa, xp = _d(x)
    Xa, Xp = _d(X)
    out = np.zeros((4, 7))
    n = lib().plo_p3p_lambdatwist(xp, Xp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out[:n]


def relpose_5pt_E(x1, x2):
    a, ap = _d(x1)
    b, bp = _d(x2)
    out = np.zeros((10, 9))
    n = lib().plo_relpose_5pt_E(ap, bp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out[:n].reshape(n, 3, 3).transpose(0, 2, 1)  # column-major -> E[r,c]


def relpose_5pt_stages(x1, x2):
    """(Nb[36], A[39], cpoly[11]) of the oracle's relpose_5pt: the nullspace basis, the 3 x 13 polynomial matrix and the
    determinant polynomial (ascending)."""
    a, ap = _d(x1)
    b, bp = _d(x2)
    out = np.zeros(86)
    lib().plo_relpose_5pt_stages(ap, bp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out[:36], out[36:75], out[75:]


def relpose_5pt(x1, x2):
    a, ap = _d(x1)
    b, bp = _d(x2)
    out = np.zeros((40, 7))
    n = lib().plo_relpose_5pt(ap, bp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out[:n]


def relpose_7pt(x1, x2):
    a, ap = _d(x1)
    b, bp = _d(x2)
    out = np.zeros((3, 9))
    n = lib().plo_relpose_7pt(ap, bp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out[:n].reshape(n, 3, 3).transpose(0, 2, 1)


def essential_matrix_8pt(x1, x2):
    """solvers/relpose_8pt.cc essential_matrix_8pt: n >= 8 unit bearing pairs -> E[r,c]."""
    a, ap = _d(x1)
    b, bp = _d(x2)
    out = np.zeros(9)
    lib().plo_essential_matrix_8pt(ap, bp, C.c_uint64(len(a)), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out.reshape(3, 3).T.copy()


def relpose_8pt(x1, x2):
    a, ap = _d(x1)
    b, bp = _d(x2)
    out = np.zeros((4, 7))
    n = lib().plo_relpose_8pt(ap, bp, C.c_uint64(len(a)), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out[:n]


def homography_4pt(x1, x2, check_cheirality=True):
    a, ap = _d(x1)
    b, bp = _d(x2)
    out = np.zeros(9)
    n = lib().plo_homography_4pt(ap, bp, out.ctypes.data_as(C.POINTER(C.c_double)), int(check_cheirality))
    return n, out.reshape(3, 3).T


def bisect_sturm10(c):
    ca, cp = _d(c)
    roots = np.zeros(10)
    n = lib().plo_bisect_sturm10(cp, roots.ctypes.data_as(C.POINTER(C.c_double)))
    return roots[:n]


def calculate_RFC(F):
    fa, fp = _d(np.asarray(F).T.reshape(-1))
    return bool(lib().plo_calculate_RFC(fp))


# ---- scorers / masks.  Matrices are passed as F[r,c] numpy arrays; poses as 7-vectors -----------
def _cm(M):
    return _d(np.asarray(M, dtype=np.float64).T.reshape(-1))


def score(kind, model, a, b, sq_thr):
    n = len(a)
    aa, ap = _d(a)
    ba, bp = _d(b)
    cnt = C.c_uint64(0)
    if kind == "pnp":
        m, mp = _d(model)
        s = lib().plo_score_pnp(mp, ap, bp, C.c_uint64(n), C.c_double(sq_thr), C.byref(cnt))
    elif kind == "relpose":
        m, mp = _d(model)
        s = lib().plo_score_relpose(mp, ap, bp, C.c_uint64(n), C.c_double(sq_thr), C.byref(cnt))
    elif kind == "fundamental":
        m, mp = _cm(model)
        s = lib().plo_score_fundamental(mp, ap, bp, C.c_uint64(n), C.c_double(sq_thr), C.byref(cnt))
    else:
        m, mp = _cm(model)
        s = lib().plo_score_homography(mp, ap, bp, C.c_uint64(n), C.c_double(sq_thr), C.byref(cnt))
    return s, cnt.value


def inliers(kind, model, a, b, sq_thr):
    n = len(a)
    aa, ap = _d(a)
    ba, bp = _d(b)
    mask, mkp = _mask(n)
    if kind == "pnp":
        m, mp = _d(model)
        lib().plo_inliers_pnp(mp, ap, bp, C.c_uint64(n), C.c_double(sq_thr), mkp)
    elif kind == "relpose":
        m, mp = _d(model)
        lib().plo_inliers_relpose(mp, ap, bp, C.c_uint64(n), C.c_double(sq_thr), mkp)
    elif kind == "fundamental":
        m, mp = _cm(model)
        lib().plo_inliers_fundamental(mp, ap, bp, C.c_uint64(n), C.c_double(sq_thr), mkp)
    else:
        m, mp = _cm(model)
        lib().plo_inliers_homography(mp, ap, bp, C.c_uint64(n), C.c_double(sq_thr), mkp)
    return mask


# ---- refiners -----------------------------------------------------------------------------------
def refine(kind, model, a, b, bopt):
    n = len(a)
    aa, ap = _d(a)
    ba, bp = _d(b)
    bs = np.zeros(7)
    bsp = bs.ctypes.data_as(C.POINTER(C.c_double))
    if kind in ("pnp", "relpose"):
        m = np.array(model, dtype=np.float64).copy()
        mp = m.ctypes.data_as(C.POINTER(C.c_double))
        fn = lib().plo_bundle_adjust if kind == "pnp" else lib().plo_refine_relpose
        fn(ap, bp, C.c_uint64(n), mp, C.byref(bopt), bsp)
        return m, bs
    m = np.ascontiguousarray(np.asarray(model, dtype=np.float64).T.reshape(-1)).copy()
    mp = m.ctypes.data_as(C.POINTER(C.c_double))
    fn = lib().plo_refine_fundamental if kind == "fundamental" else lib().plo_refine_homography
    fn(ap, bp, C.c_uint64(n), mp, C.byref(bopt), bsp)
    return m.reshape(3, 3).T.copy(), bs


# ---- RANSAC drivers -----------------------------------------------------------------------------
def ransac(kind, a, b, ropt, max_error, init=None, rfc=False):
    """Returns dict(model, inliers, stats, counters).  kind in pnp|relpose|fundamental|homography."""
    n = len(a)
    aa, ap = _d(a)
    ba, bp = _d(b)
    mask, mkp = _mask(n)
    st, cn = RansacStats(), Counters()
    if kind in ("pnp", "relpose"):
        m = np.array([1, 0, 0, 0, 0, 0, 0] if init is None else init, dtype=np.float64)
        mp = m.ctypes.data_as(C.POINTER(C.c_double))
        fn = lib().plo_ransac_pnp if kind == "pnp" else lib().plo_ransac_relpose
        fn(ap, bp, C.c_uint64(n), C.byref(ropt), C.c_double(max_error), mp, mkp, C.byref(st), C.byref(cn))
        model = m
    else:
        m0 = np.eye(3) if init is None else np.asarray(init, dtype=np.float64)
        m = np.ascontiguousarray(m0.T.reshape(-1)).copy()
        mp = m.ctypes.data_as(C.POINTER(C.c_double))
        if kind == "fundamental":
            lib().plo_ransac_fundamental(ap, bp, C.c_uint64(n), C.byref(ropt), C.c_double(max_error), int(rfc), mp,
                                         mkp, C.byref(st), C.byref(cn))
        else:
            lib().plo_ransac_homography(ap, bp, C.c_uint64(n), C.byref(ropt), C.c_double(max_error), mp, mkp,
                                        C.byref(st), C.byref(cn))
        model = m.reshape(3, 3).T.copy()
    return {"model": model, "inliers": mask, "stats": st.as_dict(), "counters": cn.as_dict()}


CAMERA_IDS = {"NULL": -1, "SIMPLE_PINHOLE": 0, "PINHOLE": 1, "SIMPLE_RADIAL": 2, "RADIAL": 3, "OPENCV": 4}


def _cam(spec):
    """None -> NULL camera; ("MODEL", params); a bare 4-sequence means PINHOLE (fx, fy, cx, cy)."""
    out = np.zeros(9, dtype=np.float64)
    if spec is None:
        out[0] = -1
    elif isinstance(spec[0], str):
        out[0] = CAMERA_IDS[spec[0].upper()]
        out[1:1 + len(spec[1])] = spec[1]
    else:
        out[0] = 1
        out[1:5] = spec
    return out, out.ctypes.data_as(C.POINTER(C.c_double))


def camera_unproject_with_jac(cam, xp):
    """-> (d (n,3), M (n,3,2)) of Camera::unproject_with_jac."""
    c, cp = _cam(cam)
    xa, xpp = _d(xp)
    n = len(xp)
    out = np.zeros((n, 9))
    lib().plo_camera_unproject_with_jac(cp, xpp, C.c_uint64(n), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out[:, :3].copy(), out[:, 3:].reshape(n, 3, 2).copy()


def camera_unproject2(cam, xp):
    c, cp = _cam(cam)
    xa, xpp = _d(xp)
    n = len(xp)
    out = np.zeros((n, 2))
    lib().plo_camera_unproject2(cp, xpp, C.c_uint64(n), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def camera_project_with_jac(cam, X):
    """-> (xp (n,2) from project_with_jac, J (n,2,3), xp (n,2) from project)."""
    c, cp = _cam(cam)
    Xa, Xp = _d(X)
    n = len(X)
    out = np.zeros((n, 8))
    po = np.zeros((n, 2))
    lib().plo_camera_project_with_jac(cp, Xp, C.c_uint64(n), out.ctypes.data_as(C.POINTER(C.c_double)),
                                      po.ctypes.data_as(C.POINTER(C.c_double)))
    return out[:, :2].copy(), out[:, 2:].reshape(n, 2, 3).copy(), po


def camera_focal(cam):
    c, cp = _cam(cam)
    lib().plo_camera_focal.restype = C.c_double
    return lib().plo_camera_focal(cp)


def score_tangent(pose, d1, d2, M1, M2, sq_thr, want_inliers=False):
    n = len(d1)
    m = np.ascontiguousarray(pose, dtype=np.float64)
    a, ap = _d(d1)
    b, bp = _d(d2)
    m1, m1p = _d(np.asarray(M1).reshape(n, 6))
    m2, m2p = _d(np.asarray(M2).reshape(n, 6))
    cnt = C.c_uint64(0)
    mask, mkp = _mask(n)
    lib().plo_score_tangent.restype = C.c_double
    s = lib().plo_score_tangent(m.ctypes.data_as(C.POINTER(C.c_double)), ap, bp, m1p, m2p, C.c_uint64(n),
                                C.c_double(sq_thr), C.byref(cnt), mkp if want_inliers else None)
    return (s, cnt.value, mask) if want_inliers else (s, cnt.value)


def refine_relpose_tangent(pose, d1, d2, M1, M2, bopt):
    n = len(d1)
    m = np.ascontiguousarray(pose, dtype=np.float64).copy()
    a, ap = _d(d1)
    b, bp = _d(d2)
    m1, m1p = _d(np.asarray(M1).reshape(n, 6))
    m2, m2p = _d(np.asarray(M2).reshape(n, 6))
    bs = np.zeros(7)
    lib().plo_refine_relpose_tangent(ap, bp, m1p, m2p, C.c_uint64(n), C.byref(bopt),
                                     m.ctypes.data_as(C.POINTER(C.c_double)),
                                     bs.ctypes.data_as(C.POINTER(C.c_double)))
    return m, bs


def ransac_relpose_cameras(x1, x2, cam1, cam2, ropt, max_error):
    """ransac_relpose with cameras (tangent Sampson error, ransac.cc:155-168); points in (scaled) pixels."""
    n = len(x1)
    aa, ap = _d(x1)
    ba, bp = _d(x2)
    c1, c1p = _cam(cam1)
    c2, c2p = _cam(cam2)
    mask, mkp = _mask(n)
    st, cn = RansacStats(), Counters()
    m = np.array([1, 0, 0, 0, 0, 0, 0], dtype=np.float64)
    lib().plo_ransac_relpose_cameras(ap, bp, C.c_uint64(n), c1p, c2p, C.byref(ropt), C.c_double(max_error),
                                     m.ctypes.data_as(C.POINTER(C.c_double)), mkp, C.byref(st), C.byref(cn))
    return {"model": m, "inliers": mask, "stats": st.as_dict(), "counters": cn.as_dict()}


def estimate(kind, a, b, ropt, bopt, max_error, cam1=None, cam2=None, init=None, rfc=False, tangent_sampson=False):
    n = len(a)
    aa, ap = _d(a)
    ba, bp = _d(b)
    mask, mkp = _mask(n)
    st, cn = RansacStats(), Counters()
    c1, c1p = _cam(cam1)
    c2, c2p = _cam(cam2)
    if kind in ("pnp", "relpose"):
        m = np.array([1, 0, 0, 0, 0, 0, 0] if init is None else init, dtype=np.float64)
        mp = m.ctypes.data_as(C.POINTER(C.c_double))
        if kind == "pnp":
            lib().plo_estimate_absolute_pose(ap, bp, C.c_uint64(n), C.byref(ropt), C.byref(bopt),
                                             C.c_double(max_error), c1p, mp, mkp, C.byref(st), C.byref(cn))
        else:
            lib().plo_estimate_relative_pose(ap, bp, C.c_uint64(n), c1p, c2p, C.byref(ropt), C.byref(bopt),
                                             C.c_double(max_error), int(tangent_sampson), mp, mkp, C.byref(st),
                                             C.byref(cn))
        model = m
    else:
        m0 = np.eye(3) if init is None else np.asarray(init, dtype=np.float64)
        m = np.ascontiguousarray(m0.T.reshape(-1)).copy()
        mp = m.ctypes.data_as(C.POINTER(C.c_double))
        if kind == "fundamental":
            lib().plo_estimate_fundamental(ap, bp, C.c_uint64(n), C.byref(ropt), C.byref(bopt),
                                           C.c_double(max_error), int(rfc), mp, mkp, C.byref(st), C.byref(cn))
        else:
            lib().plo_estimate_homography(ap, bp, C.c_uint64(n), C.byref(ropt), C.byref(bopt),
                                          C.c_double(max_error), mp, mkp, C.byref(st), C.byref(cn))
        model = m.reshape(3, 3).T.copy()
    return {"model": model, "inliers": mask, "stats": st.as_dict(), "counters": cn.as_dict()}


_native = None


def native_lib():
    """The same oracle built with -march=native for the CPU it is built on (`make -C oracle native`, SURVEY §8d asks for
    this variant beside the reference-flags build); -ffp-contract=off is kept, so results are identical, only faster."""
    global _native
    if _native is None:
        subprocess.check_call(["make", "-C", _HERE, "-s", "native"])
        _native = C.CDLL(os.path.join(_HERE, "_build", "libplo_native.so"))
        _native.plo_ransac_relpose_batch_mt.restype = C.c_double
    return _native


def ransac_relpose_batch_mt(x1_list, x2_list, ropts, max_errors, threads, native=False):
    """Many relpose problems, one problem per thread at a time.  Returns (seconds, poses, stats, counters)."""
    count = len(x1_list)
    off = np.zeros(count + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(a) for a in x1_list])
    x1 = np.ascontiguousarray(np.concatenate(x1_list), dtype=np.float64)
    x2 = np.ascontiguousarray(np.concatenate(x2_list), dtype=np.float64)
    opts = (RansacOpt * count)(*ropts)
    me = np.ascontiguousarray(max_errors, dtype=np.float64)
    poses = np.zeros((count, 7))
    poses[:, 0] = 1
    stats = (RansacStats * count)()
    cnts = (Counters * count)()
    sec = (native_lib() if native else lib()).plo_ransac_relpose_batch_mt(
        x1.ctypes.data_as(C.POINTER(C.c_double)), x2.ctypes.data_as(C.POINTER(C.c_double)),
        off.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_uint64(count), opts,
        me.ctypes.data_as(C.POINTER(C.c_double)), int(threads), poses.ctypes.data_as(C.POINTER(C.c_double)),
        stats, cnts)
    return sec, poses, [s.as_dict() for s in stats], [c.as_dict() for c in cnts]


# ---- oracle/_ref: the reference's own sampler and loop templates (built by `make -C oracle ref`) --------------------
_REF_PATH = os.path.join(_HERE, "_ref", "libplref.so")
_ref = None


def ref_available(build_if_possible=True):
    """True when oracle/_ref/libplref.so exists (it is built here, where /root/reference is mounted, and travels with
    the tree; on a box without the reference sources it can only be used if already built)."""
    if not os.path.exists(_REF_PATH) and build_if_possible and os.path.isdir("/root/reference/PoseLib"):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return os.path.exists(_REF_PATH)


def ref_lib():
    global _ref
    if _ref is None:
        if not ref_available():
            raise RuntimeError("oracle/_ref/libplref.so is not built (needs /root/reference)")
        # lazy binding: the reference's utils.cc is compiled whole, its Eigen-dependent functions reference symbols the
        # Eigen stand-in only declares; they are never called
        _ref = C.CDLL(_REF_PATH, mode=os.RTLD_LAZY)
        _ref.plref_score_fundamental.restype = C.c_double
        _ref.plref_score_homography.restype = C.c_double
        _ref.plref_all_inlier_sample_probability.restype = C.c_double
        _ref.plref_compute_dynamic_max_iter.restype = C.c_uint64
    return _ref


def ref_random_ints(seed, n):
    out = np.zeros(n, dtype=np.int32)
    ref_lib().plref_random_ints(C.c_uint64(seed), C.c_uint64(n), out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


def ref_sample_table(N, K, opt, iters):
    out = np.zeros((iters, K), dtype=np.uint32)
    ref_lib().plref_sample_table(C.c_uint64(N), C.c_uint64(K), C.byref(opt), C.c_uint64(iters),
                                 out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


def ref_all_inlier_sample_probability(ni, nd, k):
    return ref_lib().plref_all_inlier_sample_probability(C.c_uint64(ni), C.c_uint64(nd), C.c_uint64(k))


def ref_compute_dynamic_max_iter(ni, nd, k, logp, mult, mn, mx):
    return ref_lib().plref_compute_dynamic_max_iter(C.c_uint64(ni), C.c_uint64(nd), C.c_uint64(k), C.c_double(logp),
                                                    C.c_double(mult), C.c_uint64(mn), C.c_uint64(mx))


def _univariate(libobj, prefix, name, coeffs, nroots):
    out = np.zeros(nroots)
    n = getattr(libobj, prefix + name)(*[C.c_double(v) for v in coeffs], out.ctypes.data_as(C.POINTER(C.c_double)))
    return n, out


def solve_quadratic_real(a, b, c, ref=False):
    return _univariate(ref_lib() if ref else lib(), "plref_" if ref else "plo_", "solve_quadratic_real", (a, b, c), 2)


def solve_cubic_single_real(c2, c1, c0, ref=False):
    return _univariate(ref_lib() if ref else lib(), "plref_" if ref else "plo_", "solve_cubic_single_real", (c2, c1, c0), 1)


def solve_cubic_real(c2, c1, c0, ref=False):
    return _univariate(ref_lib() if ref else lib(), "plref_" if ref else "plo_", "solve_cubic_real", (c2, c1, c0), 3)


def ref_bisect_sturm10(c):
    ca, cp = _d(c)
    roots = np.zeros(10)
    n = ref_lib().plref_bisect_sturm10(cp, roots.ctypes.data_as(C.POINTER(C.c_double)))
    return roots[:n].copy()


def ref_ransac_mock(nd, k, inl, opt):
    st = RansacStats()
    ref_lib().plref_ransac_mock(C.c_uint64(nd), C.c_uint64(k), C.c_uint64(inl), C.byref(opt), C.byref(st))
    return st


def ref_ransac(kind, a, b, ropt, max_error, init=None, rfc=False):
    """The REFERENCE's ransac<> / score_models<> (robust/ransac_impl.h) driving the oracle's estimators."""
    n = len(a)
    aa, ap = _d(a)
    ba, bp = _d(b)
    mask, mkp = _mask(n)
    st = RansacStats()
    if kind in ("pnp", "relpose"):
        m = np.array([1, 0, 0, 0, 0, 0, 0] if init is None else init, dtype=np.float64)
    else:
        m0 = np.eye(3) if init is None else np.asarray(init, dtype=np.float64)
        m = np.ascontiguousarray(m0.T.reshape(-1)).copy()
    ref_lib().plref_ransac({"pnp": 0, "relpose": 1, "fundamental": 2, "homography": 3}[kind], ap, bp, C.c_uint64(n),
                           C.byref(ropt), C.c_double(max_error), int(rfc), m.ctypes.data_as(C.POINTER(C.c_double)), mkp,
                           C.byref(st))
    model = m if kind in ("pnp", "relpose") else m.reshape(3, 3).T.copy()
    return {"model": model, "inliers": mask, "stats": st.as_dict()}


def p3p_root2real(b, c, ref=False):
    r = np.zeros(2)
    L = ref_lib() if ref else lib()
    ok = getattr(L, ("plref_" if ref else "plo_") + "p3p_root2real")(C.c_double(b), C.c_double(c), r.ctypes.data_as(C.POINTER(C.c_double)))
    return ok, r


def p3p_refine_lambda(l, a12, a13, a23, b12, b13, b23, ref=False):
    out = np.ascontiguousarray(l, dtype=np.float64).copy()
    L = ref_lib() if ref else lib()
    getattr(L, ("plref_" if ref else "plo_") + "p3p_refine_lambda")(out.ctypes.data_as(C.POINTER(C.c_double)),
                                                                    *[C.c_double(v) for v in (a12, a13, a23, b12, b13, b23)])
    return out


def ref_score(kind, M, x1, x2, sq_thr, want_inliers=False):
    """compute_sampson_msac_score(F) / compute_homography_msac_score (+ get_inliers / get_homography_inliers) of the
    reference's robust/utils.cc (oracle/_ref).  kind: "fundamental" | "homography"."""
    n = len(x1)
    m, mp = _cm(M)
    a, ap = _d(x1)
    b, bp = _d(x2)
    cnt = C.c_uint64(0)
    fn = ref_lib().plref_score_fundamental if kind == "fundamental" else ref_lib().plref_score_homography
    s = fn(mp, ap, bp, C.c_uint64(n), C.c_double(sq_thr), C.byref(cnt))
    if not want_inliers:
        return s, cnt.value
    mask, mkp = _mask(n)
    if kind == "fundamental":
        ref_lib().plref_inliers_fundamental(mp, ap, bp, C.c_uint64(n), C.c_double(sq_thr), mkp)
    else:
        ref_lib().plref_inliers_homography(mp, ap, bp, C.c_uint64(n), C.c_double(sq_thr), mkp)
    return s, cnt.value, mask


def ref_calculate_RFC(F):
    m, mp = _cm(F)
    return bool(ref_lib().plref_calculate_RFC(mp))


def undistort_poly(k1, k2, two, rd, ref=False):
    L = ref_lib() if ref else lib()
    fn = getattr(L, ("plref_" if ref else "plo_") + "undistort_poly")
    fn.restype = C.c_double
    return fn(C.c_double(k1), C.c_double(k2), int(two), C.c_double(rd))


def opencv_distortion(d4, x2, with_jac=False, ref=False):
    L = ref_lib() if ref else lib()
    d, dp = _d(d4)
    x, xp = _d(x2)
    out, jac = np.zeros(2), np.zeros(4)
    getattr(L, ("plref_" if ref else "plo_") + "opencv_distortion")(dp, xp, out.ctypes.data_as(C.POINTER(C.c_double)),
                                                                    jac.ctypes.data_as(C.POINTER(C.c_double)) if with_jac else None)
    return (out, jac.reshape(2, 2)) if with_jac else out


def ref_camera_project(cam, X, with_jac=False):
    """Camera::project (models 0, 1, 2, 4) / Camera::project_with_jac (models 0, 1) of the reference (oracle/_ref)."""
    mid, params = CAMERA_IDS[cam[0].upper()], np.ascontiguousarray(cam[1], dtype=np.float64)
    Xa, Xp = _d(X)
    n = len(X)
    pp = params.ctypes.data_as(C.POINTER(C.c_double))
    if with_jac:
        out = np.zeros((n, 8))
        ref_lib().plref_camera_project_with_jac(mid, pp, len(params), Xp, C.c_uint64(n), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out[:, :2].copy(), out[:, 2:].reshape(n, 2, 3).copy()
    out = np.zeros((n, 2))
    ref_lib().plref_camera_project(mid, pp, len(params), Xp, C.c_uint64(n), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def ref_camera_focal(cam):
    params = np.ascontiguousarray(cam[1], dtype=np.float64)
    ref_lib().plref_camera_focal.restype = C.c_double
    return ref_lib().plref_camera_focal(CAMERA_IDS[cam[0].upper()], params.ctypes.data_as(C.POINTER(C.c_double)), len(params))


def camera_rescale(cam, scale, ref=False):
    mid = CAMERA_IDS[cam[0].upper()]
    if ref:
        params = np.ascontiguousarray(cam[1], dtype=np.float64).copy()
        ref_lib().plref_camera_rescale(mid, params.ctypes.data_as(C.POINTER(C.c_double)), len(params), C.c_double(scale))
        return params
    p8 = np.zeros(8)
    p8[:len(cam[1])] = cam[1]
    lib().plo_camera_rescale(mid, p8.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(scale))
    return p8[:len(cam[1])]


# ---- oracle/_ref/libplref2.so: the reference's own sources for the whole path on mini-Eigen ------------------------
# (oracle/ref/ref2_capi.cc).  Its entry points mirror the oracle's (plr2_* for plo_*), so every wrapper of this module
# can be pointed at it:   with P.reference_sources(): P.estimate(...)
# libplref2_alt.so is the same build with mini-Eigen's reductions in packet / tree order (sensitivity study).
_REF2_PATHS = {"std": os.path.join(_HERE, "_ref", "libplref2.so"), "alt": os.path.join(_HERE, "_ref", "libplref2_alt.so")}
_REF2_TARGETS = {"std": "ref2", "alt": "ref2alt"}
_ref2 = {}
_use_ref2 = None


def ref2_available(build_if_possible=True, variant="std"):
    path = _REF2_PATHS[variant]
    if not os.path.exists(path) and build_if_possible and os.path.isdir("/root/reference/PoseLib"):
        with open(os.devnull, "w") as quiet:
            subprocess.check_call(["make", "-C", _HERE, "-s", _REF2_TARGETS[variant]], stdout=quiet, stderr=quiet)
    return os.path.exists(path)


def ref2_lib(variant="std"):
    if variant not in _ref2:
        if not ref2_available(variant=variant):
            raise RuntimeError(f"oracle/_ref/{os.path.basename(_REF2_PATHS[variant])} is not built (needs /root/reference)")
        # lazy binding: the estimators reference out-of-scope solvers (p4pf, gp3p, ...) that are not compiled in
        lib2 = C.CDLL(_REF2_PATHS[variant], mode=os.RTLD_LAZY)
        for n in _DOUBLE_RETURNS:
            getattr(lib2, "plr2_" + n[4:]).restype = C.c_double
        _ref2[variant] = lib2
    return _ref2[variant]


class _Ref2Proxy:
    def __getattr__(self, name):
        if not name.startswith("plo_"):
            raise AttributeError(name)
        return getattr(ref2_lib(_use_ref2), "plr2_" + name[4:])


class reference_sources:
    """Context manager: the wrappers of this module call the reference's sources (libplref2.so; alt=True: the build with
    packet / tree shaped reductions) instead of the oracle."""

    def __init__(self, alt=False):
        self._variant = "alt" if alt else "std"

    def __enter__(self):
        global _use_ref2
        self._prev = _use_ref2
        _use_ref2 = self._variant
        return self

    def __exit__(self, *exc):
        global _use_ref2
        _use_ref2 = self._prev
        return False


def set_reference_order(enable, det_terms=None):
    """Test hook: switch the oracle's 5- / 7-point solvers (and the tangent refiner's norm) to the reference's operation
    order.  det_terms: rows (k, sign, r0, c0, r1, c1, r2, c2) of the 5-point determinant expansion in the reference's order
    (parsed from the reference's source by the caller; nothing of it is stored here).  Raises if the table is not the
    complete term set.  Always addresses the oracle library, also inside `with reference_sources()`."""
    global _use_ref2
    saved, _use_ref2 = _use_ref2, None
    try:
        oracle = lib()
    finally:
        _use_ref2 = saved
    if not enable:
        oracle.plo_set_reference_order(0, None, 0)
        return
    t = np.ascontiguousarray(det_terms, dtype=np.int32).reshape(-1, 8)
    rc = oracle.plo_set_reference_order(1, t.ctypes.data_as(C.POINTER(C.c_int32)), len(t))
    if rc != 0:
        raise ValueError(f"reference-order table rejected (code {rc})")
