// ORACLE — TEST INFRASTRUCTURE ONLY.
// oracle/_ref/libplref2.so: the reference's OWN sources for the whole hot path, unmodified and compiled where they lie
// under /root/reference, running on top of "mini-Eigen" (oracle/ref/mini/Eigen/Core) because Eigen3 is not installed:
//   PoseLib/robust.cc                          estimate_absolute_pose / relative_pose / fundamental / homography
//   PoseLib/robust/ransac.cc, ransac_impl.h    ransac_pnp / relpose / fundamental / homography, LO-RANSAC loop
//   PoseLib/robust/estimators/*.cc             the estimators (sampling, minimal solver calls, scoring, LO refits)
//   PoseLib/robust/bundle.cc, optim/*.h        the LM refiners, Jacobian accumulators, robust losses
//   PoseLib/robust/utils.cc                    all scorers and inlier masks
//   PoseLib/solvers/{p3p,relpose_5pt,relpose_7pt,homography_4pt,relpose_8pt}.cc, misc/{essential,univariate,camera_models}.cc
// WHAT A MATCH PROVES: mini-Eigen implements every Eigen operation with the oracle's restatement of it
// (oracle/plo_math.h), so `reference sources + mini-Eigen == oracle` checks that the oracle transcribes PoseLib's
// LOGIC faithfully — every formula, branch, loop, call order and sign convention of the files above — but NOT that the
// shared linear-algebra primitives equal real Eigen's bit for bit (that remains unpinned).
// The entry points mirror oracle/plo_capi.cc one to one (same argument lists, prefix plr2_ instead of plo_), so the
// tests drive both through the same Python wrappers.
#include "PoseLib/misc/camera_models.h"
#include "PoseLib/misc/essential.h"
#include "PoseLib/robust.h"
#include "PoseLib/robust/bundle.h"
#include "PoseLib/robust/ransac.h"
#include "PoseLib/robust/utils.h"
#include "PoseLib/solvers/homography_4pt.h"
#include "PoseLib/solvers/p3p.h"
#include "PoseLib/solvers/p3p_lambdatwist.h"
#include "PoseLib/solvers/relpose_5pt.h"
#include "PoseLib/solvers/relpose_7pt.h"
#include "PoseLib/solvers/relpose_8pt.h"

#include <cstring>

using namespace poselib;

extern "C" {
struct plo_ransac_opt { // layouts of oracle/plo_capi.cc
    uint64_t max_iterations, min_iterations;
    double dyn_num_trials_mult, success_prob;
    uint64_t seed;
    int32_t progressive_sampling, score_initial_model;
    uint64_t max_prosac_iterations;
};
struct plo_ransac_stats {
    uint64_t refinements, iterations, num_inliers;
    double inlier_ratio, model_score;
};
struct plo_bundle_opt {
    uint64_t max_iterations;
    int32_t loss_type, pad;
    double loss_scale, gradient_tol, step_tol, relative_cost_tol, initial_lambda, min_lambda, max_lambda;
};
}

namespace {
RansacOptions cvt(const plo_ransac_opt *o) {
    RansacOptions r;
    r.max_iterations = o->max_iterations;
    r.min_iterations = o->min_iterations;
    r.dyn_num_trials_mult = o->dyn_num_trials_mult;
    r.success_prob = o->success_prob;
    r.seed = o->seed;
    r.progressive_sampling = o->progressive_sampling != 0;
    r.max_prosac_iterations = o->max_prosac_iterations;
    r.score_initial_model = o->score_initial_model != 0;
    return r;
}
BundleOptions cvt(const plo_bundle_opt *o) {
    BundleOptions b;
    b.max_iterations = o->max_iterations;
    b.loss_type = (BundleOptions::LossType)o->loss_type;
    b.loss_scale = o->loss_scale;
    b.gradient_tol = o->gradient_tol;
    b.step_tol = o->step_tol;
    b.relative_cost_tol = o->relative_cost_tol;
    b.initial_lambda = o->initial_lambda;
    b.min_lambda = o->min_lambda;
    b.max_lambda = o->max_lambda;
    return b;
}
void put(const RansacStats &s, plo_ransac_stats *o) {
    if (!o) return;
    o->refinements = s.refinements;
    o->iterations = s.iterations;
    o->num_inliers = s.num_inliers;
    o->inlier_ratio = s.inlier_ratio;
    o->model_score = s.model_score;
}
void put_bs(const BundleStats &s, double *o) {
    if (!o) return;
    o[0] = (double)s.iterations; o[1] = s.initial_cost; o[2] = s.cost; o[3] = s.lambda;
    o[4] = (double)s.invalid_steps; o[5] = s.step_norm; o[6] = s.grad_norm;
}
std::vector<Point2D> v2(const double *p, size_t n) {
    std::vector<Point2D> v(n);
    for (size_t i = 0; i < n; ++i) { v[i](0) = p[2 * i]; v[i](1) = p[2 * i + 1]; }
    return v;
}
std::vector<Point3D> v3(const double *p, size_t n) {
    std::vector<Point3D> v(n);
    for (size_t i = 0; i < n; ++i) { v[i](0) = p[3 * i]; v[i](1) = p[3 * i + 1]; v[i](2) = p[3 * i + 2]; }
    return v;
}
std::vector<Eigen::Matrix<double, 3, 2>> m32(const double *M, size_t n) {
    std::vector<Eigen::Matrix<double, 3, 2>> r(n);
    for (size_t k = 0; k < n; ++k)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j) r[k](i, j) = M[6 * k + 2 * i + j];
    return r;
}
CameraPose pose_in(const double *p) {
    CameraPose c;
    for (int i = 0; i < 4; ++i) c.q(i) = p[i];
    for (int i = 0; i < 3; ++i) c.t(i) = p[4 + i];
    return c;
}
void pose_out(const CameraPose &c, double *p) {
    for (int i = 0; i < 4; ++i) p[i] = c.q(i);
    for (int i = 0; i < 3; ++i) p[4 + i] = c.t(i);
}
Eigen::Matrix3d mat_in(const double *p) {
    Eigen::Matrix3d m;
    for (int k = 0; k < 9; ++k) m(k % 3, k / 3) = p[k];
    return m;
}
void mat_out(const Eigen::Matrix3d &m, double *p) {
    for (int k = 0; k < 9; ++k) p[k] = m(k % 3, k / 3);
}
void mask_out(const std::vector<char> &v, char *p) {
    if (p) std::copy(v.begin(), v.end(), p);
}
Camera cam_in(const double *cam9) { // model id followed by 8 parameter slots (oracle/plo_capi.cc)
    static const int np[] = {3, 4, 4, 5, 8};
    Camera c;
    c.model_id = (int)cam9[0];
    c.width = c.height = 0;
    c.params.clear();
    if (c.model_id >= 0 && c.model_id <= 4) c.params.assign(cam9 + 1, cam9 + 1 + np[c.model_id]);
    return c;
}
} // namespace

extern "C" {
// ---- minimal solvers --------------------------------------------------------------------------
int plr2_p3p(const double *x9, const double *X9, double *poses_out) {
    std::vector<CameraPose> out;
    const int n = p3p(v3(x9, 3), v3(X9, 3), &out);
    for (size_t k = 0; k < out.size(); ++k) pose_out(out[k], poses_out + 7 * k);
    return n;
}
int plr2_p3p_lambdatwist(const double *x9, const double *X9, double *poses_out) {
    std::vector<CameraPose> out;
    const int n = p3p_lambdatwist(v3(x9, 3), v3(X9, 3), &out);
    for (size_t k = 0; k < out.size(); ++k) pose_out(out[k], poses_out + 7 * k);
    return n;
}
int plr2_relpose_5pt_E(const double *x1, const double *x2, double *E_out) {
    std::vector<Eigen::Matrix3d> out;
    const int n = relpose_5pt(v3(x1, 5), v3(x2, 5), &out);
    for (size_t k = 0; k < out.size(); ++k) mat_out(out[k], E_out + 9 * k);
    return n;
}
int plr2_relpose_5pt(const double *x1, const double *x2, double *poses_out) {
    std::vector<CameraPose> out;
    const int n = relpose_5pt(v3(x1, 5), v3(x2, 5), &out);
    for (size_t k = 0; k < out.size(); ++k) pose_out(out[k], poses_out + 7 * k);
    return n;
}
int plr2_relpose_7pt(const double *x1, const double *x2, double *F_out) {
    std::vector<Eigen::Matrix3d> out;
    const int n = relpose_7pt(v3(x1, 7), v3(x2, 7), &out);
    for (size_t k = 0; k < out.size(); ++k) mat_out(out[k], F_out + 9 * k);
    return n;
}
int plr2_homography_4pt(const double *x1, const double *x2, double *H_out, int check_cheirality) {
    Eigen::Matrix3d H;
    H.setZero();
    const int n = homography_4pt(v3(x1, 4), v3(x2, 4), &H, check_cheirality != 0);
    mat_out(H, H_out);
    return n;
}
void plr2_essential_matrix_8pt(const double *x1, const double *x2, uint64_t n, double *E_out) {
    Eigen::Matrix3d E;
    essential_matrix_8pt(v3(x1, n), v3(x2, n), &E);
    mat_out(E, E_out);
}
int plr2_relpose_8pt(const double *x1, const double *x2, uint64_t n, double *poses_out) {
    CameraPoseVector out;
    const int c = relpose_8pt(v3(x1, n), v3(x2, n), &out);
    for (size_t k = 0; k < out.size(); ++k) pose_out(out[k], poses_out + 7 * k);
    return c;
}
int plr2_calculate_RFC(const double *F9) { return calculate_RFC(mat_in(F9)) ? 1 : 0; }

// ---- scorers and masks ------------------------------------------------------------------------
double plr2_score_pnp(const double *pose, const double *x, const double *X, uint64_t n, double sq_thr, uint64_t *cnt) {
    size_t c = 0;
    const double s = compute_msac_score(pose_in(pose), v2(x, n), v3(X, n), sq_thr, &c);
    *cnt = c;
    return s;
}
double plr2_score_relpose(const double *pose, const double *x1, const double *x2, uint64_t n, double sq_thr, uint64_t *cnt) {
    size_t c = 0;
    const double s = compute_sampson_msac_score(pose_in(pose), v2(x1, n), v2(x2, n), sq_thr, &c);
    *cnt = c;
    return s;
}
double plr2_score_fundamental(const double *F, const double *x1, const double *x2, uint64_t n, double sq_thr, uint64_t *cnt) {
    size_t c = 0;
    const double s = compute_sampson_msac_score(mat_in(F), v2(x1, n), v2(x2, n), sq_thr, &c);
    *cnt = c;
    return s;
}
double plr2_score_homography(const double *H, const double *x1, const double *x2, uint64_t n, double sq_thr, uint64_t *cnt) {
    size_t c = 0;
    const double s = compute_homography_msac_score(mat_in(H), v2(x1, n), v2(x2, n), sq_thr, &c);
    *cnt = c;
    return s;
}
void plr2_inliers_pnp(const double *pose, const double *x, const double *X, uint64_t n, double sq_thr, char *mask) {
    std::vector<char> m;
    get_inliers(pose_in(pose), v2(x, n), v3(X, n), sq_thr, &m);
    mask_out(m, mask);
}
int plr2_inliers_relpose(const double *pose, const double *x1, const double *x2, uint64_t n, double sq_thr, char *mask) {
    std::vector<char> m;
    const int c = get_inliers(pose_in(pose), v2(x1, n), v2(x2, n), sq_thr, &m);
    mask_out(m, mask);
    return c;
}
int plr2_inliers_fundamental(const double *F, const double *x1, const double *x2, uint64_t n, double sq_thr, char *mask) {
    std::vector<char> m;
    const int c = get_inliers(mat_in(F), v2(x1, n), v2(x2, n), sq_thr, &m);
    mask_out(m, mask);
    return c;
}
void plr2_inliers_homography(const double *H, const double *x1, const double *x2, uint64_t n, double sq_thr, char *mask) {
    std::vector<char> m;
    get_homography_inliers(mat_in(H), v2(x1, n), v2(x2, n), sq_thr, &m);
    mask_out(m, mask);
}

// ---- refiners ---------------------------------------------------------------------------------
void plr2_bundle_adjust(const double *x, const double *X, uint64_t n, double *pose, const plo_bundle_opt *opt, double *bstats7) {
    CameraPose p = pose_in(pose);
    put_bs(bundle_adjust(v2(x, n), v3(X, n), &p, cvt(opt)), bstats7);
    pose_out(p, pose);
}
void plr2_refine_relpose(const double *x1, const double *x2, uint64_t n, double *pose, const plo_bundle_opt *opt, double *bstats7) {
    CameraPose p = pose_in(pose);
    put_bs(refine_relpose(v2(x1, n), v2(x2, n), &p, cvt(opt)), bstats7);
    pose_out(p, pose);
}
void plr2_refine_fundamental(const double *x1, const double *x2, uint64_t n, double *F, const plo_bundle_opt *opt, double *bstats7) {
    Eigen::Matrix3d m = mat_in(F);
    put_bs(refine_fundamental(v2(x1, n), v2(x2, n), &m, cvt(opt)), bstats7);
    mat_out(m, F);
}
void plr2_refine_homography(const double *x1, const double *x2, uint64_t n, double *H, const plo_bundle_opt *opt, double *bstats7) {
    Eigen::Matrix3d m = mat_in(H);
    put_bs(refine_homography(v2(x1, n), v2(x2, n), &m, cvt(opt)), bstats7);
    mat_out(m, H);
}

// ---- RANSAC drivers (the trailing counters pointer of the oracle's API is accepted and ignored) -
void plr2_ransac_pnp(const double *x, const double *X, uint64_t n, const plo_ransac_opt *opt, double max_error,
                     double *pose, char *inliers, plo_ransac_stats *stats, void *) {
    AbsolutePoseOptions o;
    o.ransac = cvt(opt);
    o.max_error = max_error;
    CameraPose p = pose_in(pose);
    std::vector<char> m;
    put(ransac_pnp(v2(x, n), v3(X, n), o, &p, &m), stats);
    pose_out(p, pose);
    mask_out(m, inliers);
}
void plr2_ransac_relpose(const double *x1, const double *x2, uint64_t n, const plo_ransac_opt *opt, double max_error,
                         double *pose, char *inliers, plo_ransac_stats *stats, void *) {
    RelativePoseOptions o;
    o.ransac = cvt(opt);
    o.max_error = max_error;
    CameraPose p = pose_in(pose);
    std::vector<char> m;
    put(ransac_relpose(v2(x1, n), v2(x2, n), o, &p, &m), stats);
    pose_out(p, pose);
    mask_out(m, inliers);
}
void plr2_ransac_fundamental(const double *x1, const double *x2, uint64_t n, const plo_ransac_opt *opt,
                             double max_error, int rfc, double *F, char *inliers, plo_ransac_stats *stats, void *) {
    RelativePoseOptions o;
    o.ransac = cvt(opt);
    o.max_error = max_error;
    o.real_focal_check = rfc != 0;
    Eigen::Matrix3d M = mat_in(F);
    std::vector<char> m;
    put(ransac_fundamental(v2(x1, n), v2(x2, n), o, &M, &m), stats);
    mat_out(M, F);
    mask_out(m, inliers);
}
void plr2_ransac_homography(const double *x1, const double *x2, uint64_t n, const plo_ransac_opt *opt, double max_error,
                            double *H, char *inliers, plo_ransac_stats *stats, void *) {
    HomographyOptions o;
    o.ransac = cvt(opt);
    o.max_error = max_error;
    Eigen::Matrix3d M = mat_in(H);
    std::vector<char> m;
    put(ransac_homography(v2(x1, n), v2(x2, n), o, &M, &m), stats);
    mat_out(M, H);
    mask_out(m, inliers);
}

// ---- camera models ----------------------------------------------------------------------------
void plr2_camera_unproject_with_jac(const double *cam9, const double *xp, uint64_t n, double *out) {
    const Camera c = cam_in(cam9);
    for (uint64_t k = 0; k < n; ++k) {
        Eigen::Vector2d p(xp[2 * k], xp[2 * k + 1]);
        Eigen::Vector3d d;
        Eigen::Matrix<double, 3, 2> M;
        c.unproject_with_jac(p, &d, &M);
        for (int i = 0; i < 3; ++i) out[9 * k + i] = d(i);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j) out[9 * k + 3 + 2 * i + j] = M(i, j);
    }
}
void plr2_camera_unproject2(const double *cam9, const double *xp, uint64_t n, double *out) {
    const Camera c = cam_in(cam9);
    for (uint64_t k = 0; k < n; ++k) {
        Eigen::Vector2d p(xp[2 * k], xp[2 * k + 1]), r;
        c.unproject(p, &r);
        out[2 * k] = r(0);
        out[2 * k + 1] = r(1);
    }
}
void plr2_camera_project_with_jac(const double *cam9, const double *X, uint64_t n, double *out, double *proj_only) {
    const Camera c = cam_in(cam9);
    for (uint64_t k = 0; k < n; ++k) {
        const Eigen::Vector3d x(X[3 * k], X[3 * k + 1], X[3 * k + 2]);
        Eigen::Vector2d xp;
        Eigen::Matrix<double, 2, 3> J;
        c.project_with_jac(x, &xp, &J);
        out[8 * k] = xp(0);
        out[8 * k + 1] = xp(1);
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 3; ++j) out[8 * k + 2 + 3 * i + j] = J(i, j);
        c.project(x, &xp);
        proj_only[2 * k] = xp(0);
        proj_only[2 * k + 1] = xp(1);
    }
}
double plr2_camera_focal(const double *cam9) { return cam_in(cam9).focal(); }

// ---- tangent Sampson path ---------------------------------------------------------------------
double plr2_score_tangent(const double *pose, const double *d1, const double *d2, const double *M1, const double *M2,
                          uint64_t n, double sq_threshold, uint64_t *count, char *inliers) {
    const CameraPose p = pose_in(pose);
    size_t c = 0;
    const double s = compute_tangent_sampson_msac_score(p, v3(d1, n), v3(d2, n), m32(M1, n), m32(M2, n), sq_threshold, &c);
    *count = c;
    if (inliers) {
        std::vector<char> m;
        get_tangent_sampson_inliers(p, v3(d1, n), v3(d2, n), m32(M1, n), m32(M2, n), sq_threshold, &m);
        mask_out(m, inliers);
    }
    return s;
}
void plr2_refine_relpose_tangent(const double *d1, const double *d2, const double *M1, const double *M2, uint64_t n,
                                 const plo_bundle_opt *bopt, double *pose, double *bstats7) {
    CameraPose p = pose_in(pose);
    put_bs(refine_relpose(v3(d1, n), v3(d2, n), m32(M1, n), m32(M2, n), &p, cvt(bopt)), bstats7);
    pose_out(p, pose);
}
void plr2_ransac_relpose_cameras(const double *x1, const double *x2, uint64_t n, const double *cam1_9,
                                 const double *cam2_9, const plo_ransac_opt *opt, double max_error, double *pose,
                                 char *inliers, plo_ransac_stats *stats, void *) {
    RelativePoseOptions o;
    o.ransac = cvt(opt);
    o.max_error = max_error;
    CameraPose p = pose_in(pose);
    std::vector<char> m(n, 0);
    put(ransac_relpose(v2(x1, n), v2(x2, n), cam_in(cam1_9), cam_in(cam2_9), o, &p, &m), stats);
    pose_out(p, pose);
    mask_out(m, inliers);
}

// ---- estimate_* (PoseLib/robust.cc) -----------------------------------------------------------
void plr2_estimate_absolute_pose(const double *x, const double *X, uint64_t n, const plo_ransac_opt *ropt,
                                 const plo_bundle_opt *bopt, double max_error, const double *cam9, double *pose,
                                 char *inliers, plo_ransac_stats *stats, void *) {
    AbsolutePoseOptions o;
    o.ransac = cvt(ropt);
    o.bundle = cvt(bopt);
    o.max_error = max_error;
    Image image;
    image.camera = cam_in(cam9);
    image.pose = pose_in(pose);
    std::vector<char> m(n, 0);
    put(estimate_absolute_pose(v2(x, n), v3(X, n), o, &image, &m), stats);
    pose_out(image.pose, pose);
    mask_out(m, inliers);
}
void plr2_estimate_relative_pose(const double *x1, const double *x2, uint64_t n, const double *cam1_9,
                                 const double *cam2_9, const plo_ransac_opt *ropt, const plo_bundle_opt *bopt,
                                 double max_error, int tangent_sampson, double *pose, char *inliers,
                                 plo_ransac_stats *stats, void *) {
    RelativePoseOptions o;
    o.ransac = cvt(ropt);
    o.bundle = cvt(bopt);
    o.max_error = max_error;
    o.tangent_sampson = tangent_sampson != 0;
    CameraPose p = pose_in(pose);
    std::vector<char> m(n, 0);
    put(estimate_relative_pose(v2(x1, n), v2(x2, n), cam_in(cam1_9), cam_in(cam2_9), o, &p, &m), stats);
    pose_out(p, pose);
    mask_out(m, inliers);
}
void plr2_estimate_fundamental(const double *x1, const double *x2, uint64_t n, const plo_ransac_opt *ropt,
                               const plo_bundle_opt *bopt, double max_error, int rfc, double *F, char *inliers,
                               plo_ransac_stats *stats, void *) {
    RelativePoseOptions o;
    o.ransac = cvt(ropt);
    o.bundle = cvt(bopt);
    o.max_error = max_error;
    o.real_focal_check = rfc != 0;
    Eigen::Matrix3d M = mat_in(F);
    std::vector<char> m(n, 0);
    put(estimate_fundamental(v2(x1, n), v2(x2, n), o, &M, &m), stats);
    mat_out(M, F);
    mask_out(m, inliers);
}
void plr2_estimate_homography(const double *x1, const double *x2, uint64_t n, const plo_ransac_opt *ropt,
                              const plo_bundle_opt *bopt, double max_error, double *H, char *inliers,
                              plo_ransac_stats *stats, void *) {
    HomographyOptions o;
    o.ransac = cvt(ropt);
    o.bundle = cvt(bopt);
    o.max_error = max_error;
    Eigen::Matrix3d M = mat_in(H);
    std::vector<char> m(n, 0);
    put(estimate_homography(v2(x1, n), v2(x2, n), o, &M, &m), stats);
    mat_out(M, H);
    mask_out(m, inliers);
}
} // extern "C"
