"""CPU-only: the C-ABI library loads, exports every symbol include/poselib_b200.h declares, and fails loudly
(no CPU fallback) when no CUDA device is usable.  No compute calls need a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "poselib_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from poselib_b200 import cabi
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(cabi.lib(), s), f"{s} declared in the header but not exported"
    assert sorted(cabi.EXPORTS) == syms


def test_struct_layouts_match_header():
    from poselib_b200 import cabi
    assert C.sizeof(cabi.RansacOpt) == 56
    assert C.sizeof(cabi.RansacStats) == 40
    assert C.sizeof(cabi.BundleOpt) == 72
    assert C.sizeof(cabi.Counters) == 128
    assert C.sizeof(cabi.Camera) == 80
    o = cabi.RansacOpt(1, 2, 9.0, 0.5, 77, True, True, 5)
    cabi.lib().plb_ransac_opt_default(C.byref(o))
    assert (o.max_iterations, o.min_iterations, o.dyn_num_trials_mult, o.success_prob, o.seed,
            o.progressive_sampling, o.score_initial_model, o.max_prosac_iterations) == \
        (100000, 1000, 3.0, 0.9999, 0, 0, 0, 100000)  # PoseLib/types.h:39-50 defaults
    b = cabi.BundleOpt(1, "TRIVIAL", 7.0)
    cabi.lib().plb_bundle_opt_default(C.byref(b))
    assert (b.max_iterations, b.loss_type, b.loss_scale, b.gradient_tol, b.step_tol, b.relative_cost_tol,
            b.initial_lambda, b.min_lambda, b.max_lambda) == (100, 3, 1.0, 1e-12, 1e-8, 1e-10, 1e-3, 1e-10, 1e10)


def test_argument_errors_do_not_need_a_gpu():
    from poselib_b200 import cabi
    st, cn = cabi.RansacStats(), cabi.Counters()
    rc = cabi.lib().plb_ransac_relpose(None, None, C.c_size_t(10), None, C.c_double(1.0), None, None,
                                       C.byref(st), C.byref(cn))
    assert rc == cabi.PLB_ERR_ARG
    assert cabi.lib().plb_set_mode(7) == cabi.PLB_ERR_ARG
    # unsupported camera model -> NYI (reference throws "NYI", camera_models.cc:184-185)
    x = np.zeros((8, 2))
    cam = cabi.Camera(5, (500, 500, 0, 0))  # OPENCV_FISHEYE: not one of the six models on the path
    with pytest.raises(cabi.PoseLibB200Error) as e:
        cabi.estimate("relpose", x, x, cabi.RansacOpt(), cabi.BundleOpt(), 1.0, cam, cam)
    assert e.value.code == cabi.PLB_ERR_NYI


def test_new_entry_points_validate_arguments_without_a_gpu():
    from poselib_b200 import cabi
    x = np.zeros((8, 2))
    bad = cabi.Camera(7, (1, 1, 0, 0))  # FOV: not on the path
    good = cabi.Camera("RADIAL", (1000.0, 0.0, 0.0, -0.04, 0.001))
    for fn in (lambda: cabi.ransac_relpose_cameras(x, x, bad, good, cabi.RansacOpt(), 1.0),
               lambda: cabi.ransac_relpose_cameras(x, x, good, bad, cabi.RansacOpt(), 1.0),
               lambda: cabi.refine_relpose_cameras(np.r_[1.0, np.zeros(6)], x, x, bad, good, cabi.BundleOpt()),
               lambda: cabi.estimate("relpose", x, x, cabi.RansacOpt(), cabi.BundleOpt(), 1.0, good, bad, tangent_sampson=True),
               lambda: cabi.estimate("pnp", x, np.zeros((8, 3)), cabi.RansacOpt(), cabi.BundleOpt(), 1.0, bad)):
        with pytest.raises(cabi.PoseLibB200Error) as e:
            fn()
        assert e.value.code == cabi.PLB_ERR_NYI
    st, cn = cabi.RansacStats(), cabi.Counters()
    rc = cabi.lib().plb_ransac_relpose_cameras(None, None, C.c_size_t(10), C.byref(good), C.byref(good), None,
                                               C.c_double(1.0), None, None, C.byref(st), C.byref(cn))
    assert rc == cabi.PLB_ERR_ARG
    assert cabi.lib().plb_host_sample_table(C.c_uint64(3), C.c_uint32(5), C.byref(cabi.RansacOpt()), C.c_uint64(1),
                                            None) == cabi.PLB_ERR_ARG


def test_no_silent_cpu_fallback():
    from poselib_b200 import cabi
    if cabi.device_count() > 0:
        pytest.skip("a GPU is visible")
    x = np.random.default_rng(0).normal(size=(50, 2))
    with pytest.raises(cabi.PoseLibB200Error) as e:
        cabi.ransac("relpose", x, x, cabi.RansacOpt(max_iterations=10, min_iterations=1), 1e-3)
    assert e.value.code == cabi.PLB_ERR_CUDA
    # too few / zero points return default stats without touching the device (ransac_impl.h:161-163)
    r = cabi.ransac("relpose", np.zeros((0, 2)), np.zeros((0, 2)), cabi.RansacOpt(), 1e-3)
    assert r["stats"]["iterations"] == 0 and r["stats"]["model_score"] > 1e300


def test_product_tree_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under poselib_b200/ or include/ may include, import, link or name it,
    and the built library depends on the CUDA runtime and the C/C++ runtimes only."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"oracle|plo_py|plo_[a-z]+\.(cc|h)|libplo|plr2_|plref")
    for base in ("poselib_b200", "include"):
        for d, _, files in os.walk(os.path.join(root, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc", "Makefile")):
                    text = open(os.path.join(d, f), errors="ignore").read()
                    assert not pat.search(text), os.path.join(d, f)
    lib = os.path.join(root, "poselib_b200", "libposelib_b200.so")
    needed = subprocess.run(["readelf", "-d", lib], capture_output=True, text=True).stdout
    libs = re.findall(r"\(NEEDED\)\s+Shared library: \[(.+?)\]", needed)
    assert libs and all(re.match(r"(lib(cudart|stdc\+\+|m|gcc_s|c|pthread|dl|rt)\.so|ld-linux)", n) for n in libs), libs
