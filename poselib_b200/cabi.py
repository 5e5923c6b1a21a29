"""ctypes binding of the C-ABI (include/poselib_b200.h) — the call a Python user (or the parity tests) makes.

Every function goes through `libposelib_b200.so`; there is no Python or CPU compute path here.  If the CUDA
library is missing, import fails loudly; if no GPU is usable, the compute entry points raise PoseLibB200Error.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libposelib_b200.so")

PLB_OK, PLB_ERR_CUDA, PLB_ERR_ARG, PLB_ERR_NYI = 0, -1, -2, -3
KIND = {"pnp": 0, "relpose": 1, "fundamental": 2, "homography": 3}
LOSS = {"TRIVIAL": 0, "TRUNCATED": 1, "HUBER": 2, "CAUCHY": 3}
CAMERA = {"NULL": -1, "SIMPLE_PINHOLE": 0, "PINHOLE": 1, "SIMPLE_RADIAL": 2, "RADIAL": 3, "OPENCV": 4}


class PoseLibB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"poselib_b200 error {code}: {msg}")
        self.code = code


class RansacOpt(C.Structure):
    """PoseLib/types.h:39-50"""
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
    """PoseLib/types.h:52-58"""
    _fields_ = [("refinements", C.c_uint64), ("iterations", C.c_uint64), ("num_inliers", C.c_uint64),
                ("inlier_ratio", C.c_double), ("model_score", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class BundleOpt(C.Structure):
    """PoseLib/types.h:60-95"""
    _fields_ = [("max_iterations", C.c_uint64), ("loss_type", C.c_int32), ("reserved", C.c_int32),
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
                ("lo_calls", C.c_uint64), ("lo_seconds", C.c_double), ("gpu_launches", C.c_uint64),
                ("samples_evaluated", C.c_uint64), ("gpu_seconds", C.c_double), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("models_evaluated", C.c_uint64), ("models_confirmed", C.c_uint64),
                ("gpu_seconds_score", C.c_double), ("gpu_seconds_select", C.c_double), ("gpu_seconds_lo", C.c_double),
                ("rounds", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Camera(C.Structure):
    _fields_ = [("model_id", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("reserved", C.c_int32),
                ("params", C.c_double * 8)]

    def __init__(self, model="PINHOLE", params=(1.0, 1.0, 0.0, 0.0), width=0, height=0):
        mid = CAMERA[model] if isinstance(model, str) else int(model)
        p = list(params)[:8] + [0.0] * (8 - len(params))
        super().__init__(mid, width, height, 0, (C.c_double * 8)(*p))


class Problem(C.Structure):
    _fields_ = [("kind", C.c_int32), ("real_focal_check", C.c_int32), ("n", C.c_uint64),
                ("a", C.POINTER(C.c_double)), ("b", C.POINTER(C.c_double)), ("opt", RansacOpt),
                ("max_error", C.c_double), ("model", C.c_double * 9), ("inliers", C.c_char_p),
                ("stats", RansacStats), ("counters", Counters), ("status", C.c_int32), ("resident", C.c_int32)]


EXPORTS = [
    "plb_ransac_opt_default", "plb_bundle_opt_default", "plb_last_error", "plb_device_count", "plb_set_device",
    "plb_set_mode", "plb_host_sample_table", "plb_device_sample_table", "plb_host_dynamic_max_iter", "plb_ransac_pnp", "plb_ransac_relpose", "plb_ransac_relpose_cameras", "plb_ransac_fundamental",
    "plb_ransac_homography",
    "plb_estimate_absolute_pose", "plb_estimate_relative_pose", "plb_estimate_fundamental",
    "plb_estimate_homography", "plb_p3p_batch", "plb_p3p_lambdatwist_batch", "plb_relpose_5pt_batch", "plb_relpose_5pt_poses_batch",
    "plb_relpose_7pt_batch", "plb_homography_4pt_batch", "plb_essential_matrix_8pt_batch", "plb_relpose_8pt_batch", "plb_ransac_batch", "plb_ransac_batch_multi", "plb_estimate_batch", "plb_bundle_adjust",
    "plb_refine_relpose", "plb_refine_relpose_cameras", "plb_refine_fundamental", "plb_refine_homography", "plb_resident_create",
    "plb_resident_free",
]

class EstimateProblem(C.Structure):
    _fields_ = [("kind", C.c_int32), ("real_focal_check", C.c_int32), ("tangent_sampson", C.c_int32), ("reserved", C.c_int32),
                ("n", C.c_uint64), ("a", C.POINTER(C.c_double)), ("b", C.POINTER(C.c_double)),
                ("camera1", Camera), ("camera2", Camera), ("ransac", RansacOpt), ("bundle", BundleOpt),
                ("max_error", C.c_double), ("model", C.c_double * 9), ("inliers", C.c_char_p),
                ("stats", RansacStats), ("counters", Counters), ("status", C.c_int32), ("reserved2", C.c_int32)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
        "or make -C poselib_b200/csrc).  There is no CPU fallback.")
_lib = C.CDLL(LIB_PATH)
_lib.plb_last_error.restype = C.c_char_p
_P = C.POINTER(C.c_double)


def lib():
    return _lib


def _check(rc):
    if rc != PLB_OK:
        raise PoseLibB200Error(rc, _lib.plb_last_error().decode())


def device_count():
    return _lib.plb_device_count()


def set_device(i):
    _check(_lib.plb_set_device(int(i)))


def set_mode(mode):
    _check(_lib.plb_set_mode({"exact": 0, "fast": 1}.get(mode, mode)))


def host_sample_table(n, k, ropt, iters):
    """First `iters` minimal samples of RandomSampler(n, k, opt) as the engine's host sampler draws them (no device)."""
    out = np.zeros((iters, k), dtype=np.uint32)
    _check(_lib.plb_host_sample_table(C.c_uint64(n), C.c_uint32(k), C.byref(ropt), C.c_uint64(iters),
                                      out.ctypes.data_as(C.POINTER(C.c_uint32))))
    return out


def device_sample_table(n, k, ropt, iters, round_size=4096, count=1):
    """The same table drawn by the engine's DEVICE sampler (k_sample): `count` samplers with seeds seed+j, `round_size`
    samples per launch (state carried across launches).  Returns (count, iters, k)."""
    out = np.zeros((count, iters, k), dtype=np.uint32)
    _check(_lib.plb_device_sample_table(C.c_uint64(n), C.c_uint32(k), C.byref(ropt), C.c_uint64(iters),
                                        C.c_uint64(round_size), C.c_uint32(count),
                                        out.ctypes.data_as(C.POINTER(C.c_uint32))))
    return out


def host_dynamic_max_iter(num_inliers, num_data, sample_sz, success_prob, mult, min_it, max_it):
    _lib.plb_host_dynamic_max_iter.restype = C.c_uint64
    return _lib.plb_host_dynamic_max_iter(C.c_uint64(num_inliers), C.c_uint64(num_data), C.c_uint32(sample_sz),
                                          C.c_double(success_prob), C.c_double(mult), C.c_uint64(min_it),
                                          C.c_uint64(max_it))


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_P)


def _init_model(kind, init):
    if kind in ("pnp", "relpose"):
        m = np.array([1, 0, 0, 0, 0, 0, 0] if init is None else init, dtype=np.float64)
    else:
        m0 = np.eye(3) if init is None else np.asarray(init, dtype=np.float64)
        m = np.ascontiguousarray(m0.T.reshape(-1)).copy()  # column-major
    return m


def _model_out(kind, m):
    return m if kind in ("pnp", "relpose") else m.reshape(3, 3).T.copy()


def ransac(kind, a, b, ropt, max_error, init=None, rfc=False):
    """ransac_pnp / ransac_relpose / ransac_fundamental / ransac_homography (robust/ransac.h).
    Points are calibrated / normalised; matrices are numpy [r,c].  Returns dict(model, inliers, stats, counters)."""
    n = len(a)
    aa, ap = _d(a)
    ba, bp = _d(b)
    mask = np.zeros(max(n, 1), dtype=np.int8)
    st, cn = RansacStats(), Counters()
    m = _init_model(kind, init)
    mp = m.ctypes.data_as(_P)
    mk = mask.ctypes.data_as(C.c_char_p)
    if kind == "pnp":
        rc = _lib.plb_ransac_pnp(ap, bp, C.c_size_t(n), C.byref(ropt), C.c_double(max_error), mp, mk, C.byref(st),
                                 C.byref(cn))
    elif kind == "relpose":
        rc = _lib.plb_ransac_relpose(ap, bp, C.c_size_t(n), C.byref(ropt), C.c_double(max_error), mp, mk,
                                     C.byref(st), C.byref(cn))
    elif kind == "fundamental":
        rc = _lib.plb_ransac_fundamental(ap, bp, C.c_size_t(n), C.byref(ropt), C.c_double(max_error), int(rfc), mp,
                                         mk, C.byref(st), C.byref(cn))
    else:
        rc = _lib.plb_ransac_homography(ap, bp, C.c_size_t(n), C.byref(ropt), C.c_double(max_error), mp, mk,
                                        C.byref(st), C.byref(cn))
    _check(rc)
    return {"model": _model_out(kind, m), "inliers": mask[:n], "stats": st.as_dict(), "counters": cn.as_dict()}


def ransac_relpose_cameras(x1, x2, cam1, cam2, ropt, max_error):
    """ransac_relpose with camera models (tangent Sampson error, ransac.cc:155-168); points in pixels."""
    n = len(x1)
    aa, ap = _d(x1)
    ba, bp = _d(x2)
    mask = np.zeros(max(n, 1), dtype=np.int8)
    st, cn = RansacStats(), Counters()
    m = _init_model("relpose", None)
    _check(_lib.plb_ransac_relpose_cameras(ap, bp, C.c_size_t(n), C.byref(cam1), C.byref(cam2), C.byref(ropt),
                                           C.c_double(max_error), m.ctypes.data_as(_P),
                                           mask.ctypes.data_as(C.c_char_p), C.byref(st), C.byref(cn)))
    return {"model": _model_out("relpose", m), "inliers": mask[:n], "stats": st.as_dict(), "counters": cn.as_dict()}


def refine_relpose_cameras(pose, x1, x2, cam1, cam2, bopt):
    """refine_relpose(x1, x2, ImagePair*, opt) with fixed intrinsics (bundle.cc:237-247)."""
    n = len(x1)
    aa, ap = _d(x1)
    ba, bp = _d(x2)
    m = _init_model("relpose", pose)
    bs = np.zeros(3)
    _check(_lib.plb_refine_relpose_cameras(ap, bp, C.c_size_t(n), C.byref(cam1), C.byref(cam2),
                                           m.ctypes.data_as(_P), C.byref(bopt), bs.ctypes.data_as(_P)))
    return _model_out("relpose", m), bs


def estimate(kind, a, b, ropt, bopt, max_error, cam1=None, cam2=None, init=None, rfc=False, tangent_sampson=False):
    """estimate_absolute_pose / estimate_relative_pose / estimate_fundamental / estimate_homography (robust.h)."""
    n = len(a)
    aa, ap = _d(a)
    ba, bp = _d(b)
    mask = np.zeros(max(n, 1), dtype=np.int8)
    st, cn = RansacStats(), Counters()
    m = _init_model(kind, init)
    mp = m.ctypes.data_as(_P)
    mk = mask.ctypes.data_as(C.c_char_p)
    c1 = cam1 if cam1 is not None else Camera("NULL", ())
    c2 = cam2 if cam2 is not None else Camera("NULL", ())
    if kind == "pnp":
        rc = _lib.plb_estimate_absolute_pose(ap, bp, C.c_size_t(n), C.byref(ropt), C.byref(bopt),
                                             C.c_double(max_error), C.byref(c1), mp, mk, C.byref(st), C.byref(cn))
    elif kind == "relpose":
        rc = _lib.plb_estimate_relative_pose(ap, bp, C.c_size_t(n), C.byref(c1), C.byref(c2), C.byref(ropt),
                                             C.byref(bopt), C.c_double(max_error), int(tangent_sampson), mp, mk,
                                             C.byref(st), C.byref(cn))
    elif kind == "fundamental":
        rc = _lib.plb_estimate_fundamental(ap, bp, C.c_size_t(n), C.byref(ropt), C.byref(bopt),
                                           C.c_double(max_error), int(rfc), mp, mk, C.byref(st), C.byref(cn))
    else:
        rc = _lib.plb_estimate_homography(ap, bp, C.c_size_t(n), C.byref(ropt), C.byref(bopt),
                                          C.c_double(max_error), mp, mk, C.byref(st), C.byref(cn))
    _check(rc)
    return {"model": _model_out(kind, m), "inliers": mask[:n], "stats": st.as_dict(), "counters": cn.as_dict()}


def refine(kind, model, a, b, bopt):
    """bundle_adjust / refine_relpose / refine_fundamental / refine_homography (robust/bundle.h), uniform weights.
    Returns (model, [iterations, initial_cost, cost])."""
    n = len(a)
    aa, ap = _d(a)
    ba, bp = _d(b)
    m = _init_model(kind, model)
    bs = np.zeros(3)
    fn = {"pnp": _lib.plb_bundle_adjust, "relpose": _lib.plb_refine_relpose,
          "fundamental": _lib.plb_refine_fundamental, "homography": _lib.plb_refine_homography}[kind]
    _check(fn(ap, bp, C.c_size_t(n), m.ctypes.data_as(_P), C.byref(bopt), bs.ctypes.data_as(_P)))
    return _model_out(kind, m), bs


# ---- solvers (batched; inputs [count, k, 3] unit bearings) -------------------------------------------
def _solver(fn, a, b, per_out, extra=()):
    aa, ap = _d(a)
    ba, bp = _d(b)
    count = aa.shape[0]
    out = np.zeros((count, per_out))
    n_out = np.zeros(count, dtype=np.int32)
    _check(fn(C.c_size_t(count), ap, bp, out.ctypes.data_as(_P), n_out.ctypes.data_as(C.POINTER(C.c_int32)), *extra))
    return out, n_out


def p3p_batch(x, X):
    out, n = _solver(_lib.plb_p3p_batch, x, X, 28)
    return out.reshape(-1, 4, 7), n


def p3p_lambdatwist_batch(x, X):
    out, n = _solver(_lib.plb_p3p_lambdatwist_batch, x, X, 28)
    return out.reshape(-1, 4, 7), n


def relpose_5pt_batch(x1, x2):
    out, n = _solver(_lib.plb_relpose_5pt_batch, x1, x2, 90)
    return out.reshape(-1, 10, 3, 3).transpose(0, 1, 3, 2), n  # column-major -> [r,c]


def relpose_5pt_poses_batch(x1, x2):
    out, n = _solver(_lib.plb_relpose_5pt_poses_batch, x1, x2, 280)
    return out.reshape(-1, 40, 7), n


def relpose_7pt_batch(x1, x2):
    out, n = _solver(_lib.plb_relpose_7pt_batch, x1, x2, 27)
    return out.reshape(-1, 3, 3, 3).transpose(0, 1, 3, 2), n


def homography_4pt_batch(x1, x2, check_cheirality=True):
    out, n = _solver(_lib.plb_homography_4pt_batch, x1, x2, 9, (C.c_int(int(check_cheirality)),))
    return out.reshape(-1, 3, 3).transpose(0, 2, 1), n


def essential_matrix_8pt_batch(x1, x2):
    """solvers/relpose_8pt.h essential_matrix_8pt for [count, n, 3] unit bearings (n >= 8) -> E [count, 3, 3] ([r,c])."""
    aa, ap = _d(x1)
    ba, bp = _d(x2)
    count, n = aa.shape[0], aa.shape[1]
    out = np.zeros((count, 9))
    _check(_lib.plb_essential_matrix_8pt_batch(C.c_size_t(count), C.c_size_t(n), ap, bp, out.ctypes.data_as(_P)))
    return out.reshape(-1, 3, 3).transpose(0, 2, 1)


def relpose_8pt_batch(x1, x2):
    """solvers/relpose_8pt.h relpose_8pt for [count, n, 3] unit bearings -> (poses [count, 4, 7], n_poses [count])."""
    aa, ap = _d(x1)
    ba, bp = _d(x2)
    count, n = aa.shape[0], aa.shape[1]
    out = np.zeros((count, 28))
    n_out = np.zeros(count, dtype=np.int32)
    _check(_lib.plb_relpose_8pt_batch(C.c_size_t(count), C.c_size_t(n), ap, bp, out.ctypes.data_as(_P),
                                      n_out.ctypes.data_as(C.POINTER(C.c_int32))))
    return out.reshape(-1, 4, 7), n_out


def resident_create(kind, a, b):
    """Uploads correspondences once and keeps them in HBM; returns a handle for ransac_batch(resident=...)."""
    aa, ap = _d(a)
    ba, bp = _d(b)
    h = _lib.plb_resident_create(KIND[kind], ap, bp, C.c_size_t(len(aa)))
    if h <= 0:
        _check(h)
    return h


def resident_free(h):
    _check(_lib.plb_resident_free(int(h)))


# ---- batch of problems -------------------------------------------------------------------------------
class Batch:
    """A batch of ransac_* problems marshalled ONCE into the C-ABI's plb_problem array: run() is then just the C call
    (plb_ransac_batch / plb_ransac_batch_multi), the same batch can be run again and again (bench.py, repeated
    estimation on fixed matches) without paying the Python marshalling per call.
    problems: list of dict(kind, a, b, ransac=RansacOpt, max_error, rfc=False, init=None) or dict(kind, resident, n, ...)."""

    def __init__(self, problems):
        self.count = count = len(problems)
        self.kinds = [p["kind"] for p in problems]
        self.arr = (Problem * count)()
        self.keep = []
        self._init = []
        for i, p in enumerate(problems):
            q = self.arr[i]
            if p.get("resident"):
                aa = ba = None
                npts = int(p["n"])
                q.resident = int(p["resident"])
            else:
                aa, ap = _d(p["a"])
                ba, bp = _d(p["b"])
                npts = len(aa)
                q.a, q.b = ap, bp
            mask = np.zeros(max(npts, 1), dtype=np.int8)
            self.keep.append((aa, ba, mask))
            q.kind = KIND[p["kind"]]
            q.real_focal_check = int(p.get("rfc", False))
            q.n = npts
            q.opt = p["ransac"]
            q.max_error = p["max_error"]
            m = _init_model(p["kind"], p.get("init"))
            self._init.append(m)
            for k in range(len(m)):
                q.model[k] = m[k]
            q.inliers = C.cast(mask.ctypes.data_as(C.POINTER(C.c_char)), C.c_char_p)
        self._raw = np.frombuffer(self.arr, dtype=np.uint8).reshape(count, C.sizeof(Problem))

    def run(self, streams=8, n_gpus=None):
        """n_gpus=None: plb_ransac_batch on the current device; an int: plb_ransac_batch_multi (0 = all devices)."""
        for i, m in enumerate(self._init):  # in/out models: restore the start models of a re-run (cheap: <= 9 doubles each)
            if len(m) and self.arr[i].opt.score_initial_model:
                for k in range(len(m)):
                    self.arr[i].model[k] = m[k]
        if n_gpus is None:
            _check(_lib.plb_ransac_batch(self.arr, C.c_size_t(self.count), int(streams)))
        else:
            _check(_lib.plb_ransac_batch_multi(self.arr, C.c_size_t(self.count), int(n_gpus), int(streams)))
        return self

    def _block(self, field, ctype):
        off = getattr(Problem, field).offset
        return self._raw[:, off:off + C.sizeof(ctype)]

    def counter_sums(self):
        """Sum of every plb_counters field over the batch (vectorised over the raw array)."""
        blk = self._block("counters", Counters)
        out = {}
        for name, ct in Counters._fields_:
            o = getattr(Counters, name).offset
            col = np.ascontiguousarray(blk[:, o:o + 8]).view(np.float64 if ct is C.c_double else np.uint64)
            out[name] = float(col.sum()) if ct is C.c_double else int(col.sum())
        return out

    def records(self, indices):
        """Fixed-size result records [index, iterations, refinements, num_inliers, model_score, model(9)] (sharding.pack_results)."""
        st = np.ascontiguousarray(self._block("stats", RansacStats))
        u = st.view(np.uint64).reshape(self.count, -1)
        f = st.view(np.float64).reshape(self.count, -1)
        mo = getattr(Problem, "model").offset
        model = np.ascontiguousarray(self._raw[:, mo:mo + 72]).view(np.float64).reshape(self.count, 9)
        rec = np.zeros((self.count, 14))
        rec[:, 0] = np.asarray(indices, dtype=np.float64)
        rec[:, 1], rec[:, 2], rec[:, 3], rec[:, 4] = u[:, 1], u[:, 0], u[:, 2], f[:, 4]
        rec[:, 5:] = model
        return rec

    def results(self):
        out = []
        for i, kind in enumerate(self.kinds):
            q = self.arr[i]
            m = np.array(q.model[:7] if kind in ("pnp", "relpose") else q.model[:9], dtype=np.float64)
            out.append({"model": _model_out(kind, m), "inliers": self.keep[i][2][:q.n].copy(), "stats": q.stats.as_dict(),
                        "counters": q.counters.as_dict(), "status": q.status})
        return out


def ransac_batch(problems, streams=8, n_gpus=None):
    """problems: list of dict(kind, a, b, ransac=RansacOpt, max_error, rfc=False, init=None).
    n_gpus=None: plb_ransac_batch on the current device; an int: plb_ransac_batch_multi over that many devices of this
    process (0 = all).  Returns list of dict(model, inliers, stats, counters)."""
    return Batch(problems).run(streams, n_gpus).results()


def estimate_batch(problems, streams=8, n_gpus=-1):
    """Batch form of estimate() (plb_estimate_batch).  problems: list of dict(kind, a, b, ransac=RansacOpt, bundle=BundleOpt,
    max_error [pixels], cam1=None, cam2=None, rfc=False, tangent_sampson=False, init=None).  n_gpus: -1 current device,
    0 all devices, k the first k.  Returns list of dict(model, inliers, stats, counters, status)."""
    count = len(problems)
    arr = (EstimateProblem * count)()
    keep = []
    for i, p in enumerate(problems):
        q = arr[i]
        aa, ap = _d(p["a"])
        ba, bp = _d(p["b"])
        npts = len(aa)
        mask = np.zeros(max(npts, 1), dtype=np.int8)
        keep.append((aa, ba, mask))
        q.kind = KIND[p["kind"]]
        q.real_focal_check = int(p.get("rfc", False))
        q.tangent_sampson = int(p.get("tangent_sampson", False))
        q.n = npts
        q.a, q.b = ap, bp
        q.camera1 = p.get("cam1") or Camera("NULL", ())
        q.camera2 = p.get("cam2") or Camera("NULL", ())
        q.ransac = p["ransac"]
        q.bundle = p["bundle"]
        q.max_error = p["max_error"]
        m = _init_model(p["kind"], p.get("init"))
        for k in range(len(m)):
            q.model[k] = m[k]
        q.inliers = C.cast(mask.ctypes.data_as(C.POINTER(C.c_char)), C.c_char_p)
    _check(_lib.plb_estimate_batch(arr, C.c_size_t(count), int(n_gpus), int(streams)))
    out = []
    for i, p in enumerate(problems):
        q = arr[i]
        kind = p["kind"]
        m = np.array(q.model[:7] if kind in ("pnp", "relpose") else q.model[:9], dtype=np.float64)
        out.append({"model": _model_out(kind, m), "inliers": keep[i][2][:q.n].copy(), "stats": q.stats.as_dict(),
                    "counters": q.counters.as_dict(), "status": q.status})
    return out
