// ORACLE — TEST INFRASTRUCTURE ONLY.
// oracle/_ref/libplref.so: the parts of the UNMODIFIED reference that compile in this image without Eigen, built from
// the sources where they lie under /root/reference (recipe: oracle/Makefile, target `ref`):
//   * PoseLib/robust/sampling.cc            — random_int, draw_sample, RandomSampler incl. PROSAC
//   * PoseLib/robust/ransac_impl.h          — all_inlier_sample_probability, compute_dynamic_max_iter, score_models<>,
//                                             ransac<> (header-only templates, instantiated here)
//   * PoseLib/misc/univariate.cc            — solve_quadratic_real, solve_cubic_single_real, solve_cubic_real
//   * PoseLib/misc/sturm.h                  — bisect_sturm<10> (Sturm sequence, root isolation, Ridders + Newton)
//   * PoseLib/solvers/p3p_common.h          — root2real, refine_lambda (the scalar helpers of p3p)
//   * PoseLib/robust/utils.cc               — the five functions of that file that only read matrix / point ELEMENTS
//                                             and do scalar arithmetic: compute_sampson_msac_score(F), get_inliers(F),
//                                             compute_homography_msac_score, get_homography_inliers, calculate_RFC.
//                                             (The rest of utils.cc compiles against the Eigen stand-in but references
//                                             undefined Eigen symbols and is never called; the library is therefore
//                                             loaded with lazy binding.)
//   * PoseLib/misc/camera_models.cc         — likewise the scalar parts: undistort_poly1/2 (Newton undistortion),
//                                             compute_opencv_distortion(_jac), Camera::focal / rescale, and through
//                                             Camera::project / project_with_jac the PINHOLE, SIMPLE_PINHOLE,
//                                             SIMPLE_RADIAL and OPENCV projections that only touch elements
// The templates are instantiated (a) with the MockEstimator of the reference's tests/ransac_test.cc:12-28 and (b) with
// this repository's oracle estimators (solver / scorer / refiner restatements of oracle/plo_robust.cc), so that the
// REFERENCE's loop drives them: comparing the outcome with the oracle's own loop pins the control flow of
// score_models / ransac (best-minimal bookkeeping, LO trigger, dynamic iteration count, break test, final refinement)
// to the reference's code.  The arithmetic inside the estimators remains the oracle's (parity unpinned).
#include "PoseLib/misc/sturm.h"
#include "PoseLib/misc/univariate.h"
#include "PoseLib/robust/ransac_impl.h"
#include "PoseLib/robust/sampling.h"
#include "PoseLib/robust/utils.h"
#include "PoseLib/solvers/p3p_common.h"

#include "../plo_robust.cc" // unity include: the oracle estimators live in an anonymous namespace

#include <cstring>

namespace {

poselib::RansacOptions to_ref(const plo::RansacOptions &o) {
    poselib::RansacOptions r;
    r.max_iterations = o.max_iterations;
    r.min_iterations = o.min_iterations;
    r.dyn_num_trials_mult = o.dyn_num_trials_mult;
    r.success_prob = o.success_prob;
    r.seed = o.seed;
    r.progressive_sampling = o.progressive_sampling;
    r.max_prosac_iterations = o.max_prosac_iterations;
    r.score_initial_model = o.score_initial_model;
    return r;
}

struct COpt { // layout of plo_ransac_opt (oracle/plo_capi.cc)
    uint64_t max_iterations, min_iterations;
    double dyn_num_trials_mult, success_prob;
    uint64_t seed;
    int32_t progressive_sampling, score_initial_model;
    uint64_t max_prosac_iterations;
};
plo::RansacOptions from_c(const COpt *o) {
    plo::RansacOptions r;
    r.max_iterations = o->max_iterations;
    r.min_iterations = o->min_iterations;
    r.dyn_num_trials_mult = o->dyn_num_trials_mult;
    r.success_prob = o->success_prob;
    r.seed = o->seed;
    r.progressive_sampling = o->progressive_sampling != 0;
    r.score_initial_model = o->score_initial_model != 0;
    r.max_prosac_iterations = o->max_prosac_iterations;
    return r;
}
struct CStats {
    uint64_t refinements, iterations, num_inliers;
    double inlier_ratio, model_score;
};
void put(const poselib::RansacStats &s, CStats *o) {
    o->refinements = s.refinements;
    o->iterations = s.iterations;
    o->num_inliers = s.num_inliers;
    o->inlier_ratio = s.inlier_ratio;
    o->model_score = s.model_score;
}

// tests/ransac_test.cc:12-28 (same semantics: one model per iteration, score 0, fixed inlier count)
struct RefMockEstimator {
    size_t num_data, sample_sz, inlier_count;
    void generate_models(std::vector<int> *models) const { models->push_back(0); }
    double score_model(const int &, size_t *ic) const {
        *ic = inlier_count;
        return 0.0;
    }
    void refine_model(int *) const {}
};

std::vector<plo::Vec2> v2(const double *p, uint64_t n) {
    std::vector<plo::Vec2> r(n);
    for (uint64_t k = 0; k < n; ++k) {
        r[k][0] = p[2 * k];
        r[k][1] = p[2 * k + 1];
    }
    return r;
}
std::vector<plo::Vec3> v3(const double *p, uint64_t n) {
    std::vector<plo::Vec3> r(n);
    for (uint64_t k = 0; k < n; ++k) r[k] = plo::mk3(p[3 * k], p[3 * k + 1], p[3 * k + 2]);
    return r;
}

} // namespace

extern "C" {

// reference random_int stream (sampling.cc:37-43)
void plref_random_ints(uint64_t seed, uint64_t n, int32_t *out) {
    poselib::RNG_t st = seed;
    for (uint64_t i = 0; i < n; ++i) out[i] = poselib::random_int(st);
}
// reference RandomSampler (sampling.h:49-83, sampling.cc:85-136)
void plref_sample_table(uint64_t N, uint64_t K, const COpt *opt, uint64_t iters, uint32_t *out) {
    poselib::RandomSampler smp(N, K, to_ref(from_c(opt)));
    std::vector<size_t> s(K);
    for (uint64_t i = 0; i < iters; ++i) {
        smp.generate_sample(&s);
        for (uint64_t j = 0; j < K; ++j) out[i * K + j] = (uint32_t)s[j];
    }
}
double plref_all_inlier_sample_probability(uint64_t ni, uint64_t nd, uint64_t k) {
    return poselib::detail::all_inlier_sample_probability(ni, nd, k);
}
uint64_t plref_compute_dynamic_max_iter(uint64_t ni, uint64_t nd, uint64_t k, double logp, double mult, uint64_t mn,
                                        uint64_t mx) {
    return poselib::detail::compute_dynamic_max_iter(ni, nd, k, logp, mult, mn, mx);
}
void plref_ransac_mock(uint64_t nd, uint64_t k, uint64_t inl, const COpt *opt, CStats *stats) {
    RefMockEstimator est{(size_t)nd, (size_t)k, (size_t)inl};
    int best = 0;
    put(poselib::ransac<RefMockEstimator, int>(est, to_ref(from_c(opt)), &best), stats);
}

// misc/univariate.cc and misc/sturm.h of the reference (scalar code)
int plref_solve_quadratic_real(double a, double b, double c, double *roots) {
    return poselib::univariate::solve_quadratic_real(a, b, c, roots);
}
int plref_solve_cubic_single_real(double c2, double c1, double c0, double *root) {
    return poselib::univariate::solve_cubic_single_real(c2, c1, c0, *root) ? 1 : 0;
}
int plref_solve_cubic_real(double c2, double c1, double c0, double *roots) {
    return poselib::univariate::solve_cubic_real(c2, c1, c0, roots);
}
// misc/camera_models.cc: scalar functions defined there at namespace scope (not declared in a header)
extern "C++" {
namespace poselib {
double undistort_poly1(double k1, double rd);
double undistort_poly2(double k1, double k2, double rd);
void compute_opencv_distortion(double k1, double k2, double p1, double p2, const Eigen::Vector2d &x, Eigen::Vector2d &xp);
void compute_opencv_distortion_jac(double k1, double k2, double p1, double p2, const Eigen::Vector2d &x, Eigen::Vector2d &xp,
                                   Eigen::Matrix2d &jac, Eigen::Matrix<double, 2, 4> *jacp);
} // namespace poselib
} // extern "C++"
static poselib::Camera cam_ref(int model_id, const double *params, int np) {
    poselib::Camera c;
    c.model_id = model_id;
    c.params.assign(params, params + np);
    return c;
}
double plref_undistort_poly(double k1, double k2, int two, double rd) {
    return two ? poselib::undistort_poly2(k1, k2, rd) : poselib::undistort_poly1(k1, rd);
}
// d4 = k1,k2,p1,p2; out2 = distorted point; jac4 (may be null) = row-major 2x2
void plref_opencv_distortion(const double *d4, const double *x2, double *out2, double *jac4) {
    Eigen::Vector2d x, xp;
    x(0) = x2[0];
    x(1) = x2[1];
    if (jac4) {
        Eigen::Matrix2d J;
        poselib::compute_opencv_distortion_jac(d4[0], d4[1], d4[2], d4[3], x, xp, J, nullptr);
        jac4[0] = J(0, 0); jac4[1] = J(0, 1); jac4[2] = J(1, 0); jac4[3] = J(1, 1);
    } else {
        poselib::compute_opencv_distortion(d4[0], d4[1], d4[2], d4[3], x, xp);
    }
    out2[0] = xp(0);
    out2[1] = xp(1);
}
// Camera::project for the models whose projection only touches elements (ids 0 SIMPLE_PINHOLE, 1 PINHOLE,
// 2 SIMPLE_RADIAL, 4 OPENCV); out: n x 2
void plref_camera_project(int model_id, const double *params, int np, const double *X, uint64_t n, double *out) {
    const poselib::Camera c = cam_ref(model_id, params, np);
    for (uint64_t k = 0; k < n; ++k) {
        Eigen::Vector3d x;
        x(0) = X[3 * k]; x(1) = X[3 * k + 1]; x(2) = X[3 * k + 2];
        Eigen::Vector2d xp;
        c.project(x, &xp);
        out[2 * k] = xp(0);
        out[2 * k + 1] = xp(1);
    }
}
// Camera::project_with_jac (point Jacobian, no parameter Jacobian) for the pinhole family (ids 0, 1);
// out: n x (2 + 6 row-major 2x3)
void plref_camera_project_with_jac(int model_id, const double *params, int np, const double *X, uint64_t n, double *out) {
    const poselib::Camera c = cam_ref(model_id, params, np);
    for (uint64_t k = 0; k < n; ++k) {
        Eigen::Vector3d x;
        x(0) = X[3 * k]; x(1) = X[3 * k + 1]; x(2) = X[3 * k + 2];
        Eigen::Vector2d xp;
        Eigen::Matrix<double, 2, 3> J;
        c.project_with_jac(x, &xp, &J);
        out[8 * k] = xp(0);
        out[8 * k + 1] = xp(1);
        for (int r = 0; r < 2; ++r)
            for (int q = 0; q < 3; ++q) out[8 * k + 2 + 3 * r + q] = J(r, q);
    }
}
double plref_camera_focal(int model_id, const double *params, int np) { return cam_ref(model_id, params, np).focal(); }
void plref_camera_rescale(int model_id, double *params, int np, double scale) {
    poselib::Camera c = cam_ref(model_id, params, np);
    c.rescale(scale);
    for (int i = 0; i < np; ++i) params[i] = c.params[i];
}

// robust/utils.cc: element-access-only scorers / masks / RFC of the reference.  F, H column-major 9 doubles.
static Eigen::Matrix3d mat_in(const double *m9) {
    Eigen::Matrix3d M;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) M(r, c) = m9[3 * c + r];
    return M;
}
static std::vector<poselib::Point2D> pts_in(const double *p, uint64_t n) {
    std::vector<poselib::Point2D> v(n);
    for (uint64_t k = 0; k < n; ++k) {
        v[k](0) = p[2 * k];
        v[k](1) = p[2 * k + 1];
    }
    return v;
}
double plref_score_fundamental(const double *F9, const double *x1, const double *x2, uint64_t n, double sq_thr,
                               uint64_t *count) {
    size_t c = 0;
    const double s = poselib::compute_sampson_msac_score(mat_in(F9), pts_in(x1, n), pts_in(x2, n), sq_thr, &c);
    *count = c;
    return s;
}
double plref_score_homography(const double *H9, const double *x1, const double *x2, uint64_t n, double sq_thr,
                              uint64_t *count) {
    size_t c = 0;
    const double s = poselib::compute_homography_msac_score(mat_in(H9), pts_in(x1, n), pts_in(x2, n), sq_thr, &c);
    *count = c;
    return s;
}
int plref_inliers_fundamental(const double *F9, const double *x1, const double *x2, uint64_t n, double sq_thr, char *mask) {
    std::vector<char> m;
    const int c = poselib::get_inliers(mat_in(F9), pts_in(x1, n), pts_in(x2, n), sq_thr, &m);
    std::memcpy(mask, m.data(), n);
    return c;
}
void plref_inliers_homography(const double *H9, const double *x1, const double *x2, uint64_t n, double sq_thr, char *mask) {
    std::vector<char> m;
    poselib::get_homography_inliers(mat_in(H9), pts_in(x1, n), pts_in(x2, n), sq_thr, &m);
    std::memcpy(mask, m.data(), n);
}
int plref_calculate_RFC(const double *F9) { return poselib::calculate_RFC(mat_in(F9)) ? 1 : 0; }

int plref_p3p_root2real(double b, double c, double *r) { return poselib::root2real(b, c, r[0], r[1]) ? 1 : 0; }
void plref_p3p_refine_lambda(double *l, double a12, double a13, double a23, double b12, double b13, double b23) {
    poselib::refine_lambda(l[0], l[1], l[2], a12, a13, a23, b12, b13, b23);
}
int plref_bisect_sturm10(const double *coeffs11, double *roots10) {
    return poselib::sturm::bisect_sturm<10>(coeffs11, roots10);
}

// The reference's ransac<> / score_models<> driving the oracle estimators.  kind: 0 pnp, 1 relpose, 2 fundamental,
// 3 homography.  model_inout: 7 doubles (q, t) or 9 doubles column-major.
void plref_ransac(int kind, const double *a, const double *b, uint64_t n, const COpt *copt, double max_error, int rfc,
                  double *model_inout, char *inliers, CStats *stats) {
    const plo::RansacOptions po = from_c(copt);
    const poselib::RansacOptions ro = to_ref(po);
    std::vector<char> mask(n, 0);
    if (kind == 0 || kind == 1) {
        plo::CameraPose pose;
        if (po.score_initial_model) {
            for (int i = 0; i < 4; ++i) pose.q[i] = model_inout[i];
            for (int i = 0; i < 3; ++i) pose.t[i] = model_inout[4 + i];
        }
        if (kind == 0) {
            const auto x = v2(a, n);
            const auto X = v3(b, n);
            plo::AbsolutePoseEstimator est(po, max_error, x, X, nullptr);
            put(poselib::ransac<plo::AbsolutePoseEstimator, plo::CameraPose>(est, ro, &pose), stats);
            plo::get_inliers(pose, x, X, max_error * max_error, &mask);
        } else {
            const auto x1 = v2(a, n), x2 = v2(b, n);
            plo::RelativePoseEstimator est(po, max_error, x1, x2, nullptr);
            put(poselib::ransac<plo::RelativePoseEstimator, plo::CameraPose>(est, ro, &pose), stats);
            plo::get_inliers(pose, x1, x2, max_error * max_error, &mask);
        }
        for (int i = 0; i < 4; ++i) model_inout[i] = pose.q[i];
        for (int i = 0; i < 3; ++i) model_inout[4 + i] = pose.t[i];
    } else {
        plo::Mat3 M = plo::mat3_identity();
        if (po.score_initial_model)
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) M(r, c) = model_inout[3 * c + r];
        const auto x1 = v2(a, n), x2 = v2(b, n);
        if (kind == 2) {
            plo::FundamentalEstimator est(po, max_error, rfc != 0, x1, x2, nullptr);
            put(poselib::ransac<plo::FundamentalEstimator, plo::Mat3>(est, ro, &M), stats);
            plo::get_inliers(M, x1, x2, max_error * max_error, &mask);
        } else {
            plo::HomographyEstimator est(po, max_error, x1, x2, nullptr);
            put(poselib::ransac<plo::HomographyEstimator, plo::Mat3>(est, ro, &M), stats);
            plo::get_homography_inliers(M, x1, x2, max_error * max_error, &mask);
        }
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) model_inout[3 * c + r] = M(r, c);
    }
    if (inliers) std::memcpy(inliers, mask.data(), n);
}

} // extern "C"
