// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY PARTLY PINNED: sampler, loop control flow, iteration arithmetic, univariate / p3p scalar solvers, Sturm root isolation, F / H scorers, masks, the real-focal check and the scalar camera code against the reference's own code (oracle/_ref, oracle/ref/ref_capi.cc); the transcription of PoseLib's logic for the WHOLE path (solvers, scorers, refiners, estimators, estimate_*) against the reference's own sources run on mini-Eigen (oracle/_ref/libplref2.so, oracle/ref/ref2_capi.cc, tests/test_ref_sources.py); Eigen's own arithmetic (reduction order, decompositions) is UNPINNED.
// extern "C" surface of the CPU oracle for ctypes (tests/, __graft_entry__.smoke(), bench.py cpu legs).
// 3x3 matrices cross this boundary as 9 doubles COLUMN-MAJOR (Eigen::Matrix3d layout); poses as q(wxyz)+t.
#include "plo.h"
#include <chrono>
#include <thread>
#include <atomic>

namespace plo {
extern thread_local double *g_relpose_5pt_stage_out; // plo_solvers.cc (test hook)
}
using namespace plo;

extern "C" {

struct plo_ransac_opt {
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
struct plo_counters {
    uint64_t samples, hypotheses, scored_corrs, lo_calls;
    double lo_seconds;
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
void put(const Counters &c, plo_counters *o) {
    if (!o) return;
    o->samples = c.samples;
    o->hypotheses = c.hypotheses;
    o->scored_corrs = c.scored_corrs;
    o->lo_calls = c.lo_calls;
    o->lo_seconds = c.lo_seconds;
}
std::vector<Vec2> v2(const double *p, size_t n) {
    std::vector<Vec2> v(n);
    for (size_t i = 0; i < n; ++i) { v[i][0] = p[2 * i]; v[i][1] = p[2 * i + 1]; }
    return v;
}
std::vector<Vec3> v3(const double *p, size_t n) {
    std::vector<Vec3> v(n);
    for (size_t i = 0; i < n; ++i) { v[i][0] = p[3 * i]; v[i][1] = p[3 * i + 1]; v[i][2] = p[3 * i + 2]; }
    return v;
}
CameraPose pose_in(const double *p) {
    CameraPose c;
    for (int i = 0; i < 4; ++i) c.q[i] = p[i];
    for (int i = 0; i < 3; ++i) c.t[i] = p[4 + i];
    return c;
}
void pose_out(const CameraPose &c, double *p) {
    for (int i = 0; i < 4; ++i) p[i] = c.q[i];
    for (int i = 0; i < 3; ++i) p[4 + i] = c.t[i];
}
Mat3 mat_in(const double *p) {
    Mat3 m;
    for (int k = 0; k < 9; ++k) m(k % 3, k / 3) = p[k];
    return m;
}
void mat_out(const Mat3 &m, double *p) {
    for (int k = 0; k < 9; ++k) p[k] = m(k % 3, k / 3);
}
void mask_out(const std::vector<char> &v, char *p) {
    if (p) std::copy(v.begin(), v.end(), p);
}
} // namespace

extern "C" {

// ---- sampler ---------------------------------------------------------------------------------
void plo_random_ints(uint64_t seed, int n, int32_t *out) {
    uint64_t st = seed;
    for (int i = 0; i < n; ++i) out[i] = random_int(st);
}
void plo_sample_table(uint64_t N, uint64_t K, const plo_ransac_opt *opt, uint64_t iters, uint32_t *out) {
    RandomSampler s(N, K, cvt(opt));
    std::vector<size_t> sample(K);
    for (uint64_t i = 0; i < iters; ++i) {
        s.generate_sample(&sample);
        for (uint64_t k = 0; k < K; ++k) out[i * K + k] = (uint32_t)sample[k];
    }
}
// ---- loop KATs -------------------------------------------------------------------------------
double plo_all_inlier_sample_probability(uint64_t ni, uint64_t nd, uint64_t k) {
    return all_inlier_sample_probability(ni, nd, k);
}
uint64_t plo_compute_dynamic_max_iter(uint64_t ni, uint64_t nd, uint64_t k, double logp, double mult, uint64_t mn,
                                      uint64_t mx) {
    return compute_dynamic_max_iter(ni, nd, k, logp, mult, mn, mx);
}
void plo_ransac_mock(uint64_t nd, uint64_t k, uint64_t inl, const plo_ransac_opt *opt, plo_ransac_stats *out) {
    put(ransac_mock(nd, k, inl, cvt(opt)), out);
}
// ---- solvers ---------------------------------------------------------------------------------
int plo_p3p(const double *x9, const double *X9, double *poses_out) {
    std::vector<CameraPose> out;
    int n = p3p(v3(x9, 3), v3(X9, 3), &out);
    for (int i = 0; i < n; ++i) pose_out(out[i], poses_out + 7 * i);
    return n;
}
int plo_p3p_lambdatwist(const double *x9, const double *X9, double *poses_out) {
    std::vector<CameraPose> out;
    int n = p3p_lambdatwist(v3(x9, 3), v3(X9, 3), &out);
    for (int i = 0; i < n; ++i) pose_out(out[i], poses_out + 7 * i);
    return n;
}
int plo_relpose_5pt_E(const double *x1, const double *x2, double *E_out) {
    std::vector<Mat3> out;
    int n = relpose_5pt(v3(x1, 5), v3(x2, 5), &out);
    for (int i = 0; i < n; ++i) mat_out(out[i], E_out + 9 * i);
    return n;
}
// intermediates of relpose_5pt (test hook, see plo_solvers.cc): out86 = Nb (36) | A (39) | determinant polynomial (11)
int plo_relpose_5pt_stages(const double *x1, const double *x2, double *out86) {
    std::vector<Mat3> out;
    g_relpose_5pt_stage_out = out86;
    int n = relpose_5pt(v3(x1, 5), v3(x2, 5), &out);
    g_relpose_5pt_stage_out = nullptr;
    return n;
}
int plo_relpose_5pt(const double *x1, const double *x2, double *poses_out) {
    std::vector<CameraPose> out;
    int n = relpose_5pt(v3(x1, 5), v3(x2, 5), &out);
    for (int i = 0; i < n; ++i) pose_out(out[i], poses_out + 7 * i);
    return n;
}
int plo_relpose_7pt(const double *x1, const double *x2, double *F_out) {
    std::vector<Mat3> out;
    int n = relpose_7pt(v3(x1, 7), v3(x2, 7), &out);
    for (int i = 0; i < n; ++i) mat_out(out[i], F_out + 9 * i);
    return n;
}
int plo_homography_4pt(const double *x1, const double *x2, double *H_out, int check_cheirality) {
    Mat3 H = mat3_zero();
    int n = homography_4pt(v3(x1, 4), v3(x2, 4), &H, check_cheirality != 0);
    mat_out(H, H_out);
    return n;
}
void plo_essential_matrix_8pt(const double *x1, const double *x2, uint64_t n, double *E_out) {
    Mat3 E;
    essential_matrix_8pt(v3(x1, n), v3(x2, n), &E);
    mat_out(E, E_out);
}
int plo_relpose_8pt(const double *x1, const double *x2, uint64_t n, double *poses_out) {
    std::vector<CameraPose> out;
    const int c = relpose_8pt(v3(x1, n), v3(x2, n), &out);
    for (size_t k = 0; k < out.size(); ++k) pose_out(out[k], poses_out + 7 * k);
    return c;
}
int plo_bisect_sturm10(const double *c11, double *roots) { return bisect_sturm10(c11, roots); }
int plo_solve_quadratic_real(double a, double b, double c, double *roots) { return solve_quadratic_real(a, b, c, roots); }
int plo_solve_cubic_single_real(double c2, double c1, double c0, double *root) {
    return solve_cubic_single_real(c2, c1, c0, *root) ? 1 : 0;
}
int plo_solve_cubic_real(double c2, double c1, double c0, double *roots) { return solve_cubic_real(c2, c1, c0, roots); }
int plo_calculate_RFC(const double *F9) { return calculate_RFC(mat_in(F9)) ? 1 : 0; }
// ---- scorers / masks --------------------------------------------------------------------------
double plo_score_pnp(const double *pose, const double *x, const double *X, uint64_t n, double sq_thr, uint64_t *cnt) {
    size_t c;
    double s = compute_msac_score(pose_in(pose), v2(x, n), v3(X, n), sq_thr, &c);
    *cnt = c;
    return s;
}
double plo_score_relpose(const double *pose, const double *x1, const double *x2, uint64_t n, double sq_thr,
                         uint64_t *cnt) {
    size_t c;
    double s = compute_sampson_msac_score(pose_in(pose), v2(x1, n), v2(x2, n), sq_thr, &c);
    *cnt = c;
    return s;
}
double plo_score_fundamental(const double *F, const double *x1, const double *x2, uint64_t n, double sq_thr,
                             uint64_t *cnt) {
    size_t c;
    double s = compute_sampson_msac_score(mat_in(F), v2(x1, n), v2(x2, n), sq_thr, &c);
    *cnt = c;
    return s;
}
double plo_score_homography(const double *H, const double *x1, const double *x2, uint64_t n, double sq_thr,
                            uint64_t *cnt) {
    size_t c;
    double s = compute_homography_msac_score(mat_in(H), v2(x1, n), v2(x2, n), sq_thr, &c);
    *cnt = c;
    return s;
}
void plo_inliers_pnp(const double *pose, const double *x, const double *X, uint64_t n, double sq_thr, char *mask) {
    std::vector<char> m;
    get_inliers(pose_in(pose), v2(x, n), v3(X, n), sq_thr, &m);
    mask_out(m, mask);
}
int plo_inliers_relpose(const double *pose, const double *x1, const double *x2, uint64_t n, double sq_thr, char *mask) {
    std::vector<char> m;
    int c = get_inliers(pose_in(pose), v2(x1, n), v2(x2, n), sq_thr, &m);
    mask_out(m, mask);
    return c;
}
int plo_inliers_fundamental(const double *F, const double *x1, const double *x2, uint64_t n, double sq_thr, char *mask) {
    std::vector<char> m;
    int c = get_inliers(mat_in(F), v2(x1, n), v2(x2, n), sq_thr, &m);
    mask_out(m, mask);
    return c;
}
void plo_inliers_homography(const double *H, const double *x1, const double *x2, uint64_t n, double sq_thr, char *mask) {
    std::vector<char> m;
    get_homography_inliers(mat_in(H), v2(x1, n), v2(x2, n), sq_thr, &m);
    mask_out(m, mask);
}
// ---- refiners ---------------------------------------------------------------------------------
static void put_bs(const BundleStats &s, double *o) {
    if (!o) return;
    o[0] = (double)s.iterations; o[1] = s.initial_cost; o[2] = s.cost; o[3] = s.lambda;
    o[4] = (double)s.invalid_steps; o[5] = s.step_norm; o[6] = s.grad_norm;
}
void plo_bundle_adjust(const double *x, const double *X, uint64_t n, double *pose, const plo_bundle_opt *opt,
                       double *bstats7) {
    CameraPose p = pose_in(pose);
    put_bs(bundle_adjust(v2(x, n), v3(X, n), &p, cvt(opt)), bstats7);
    pose_out(p, pose);
}
void plo_refine_relpose(const double *x1, const double *x2, uint64_t n, double *pose, const plo_bundle_opt *opt,
                        double *bstats7) {
    CameraPose p = pose_in(pose);
    put_bs(refine_relpose(v2(x1, n), v2(x2, n), &p, cvt(opt)), bstats7);
    pose_out(p, pose);
}
void plo_refine_fundamental(const double *x1, const double *x2, uint64_t n, double *F, const plo_bundle_opt *opt,
                            double *bstats7) {
    Mat3 m = mat_in(F);
    put_bs(refine_fundamental(v2(x1, n), v2(x2, n), &m, cvt(opt)), bstats7);
    mat_out(m, F);
}
void plo_refine_homography(const double *x1, const double *x2, uint64_t n, double *H, const plo_bundle_opt *opt,
                           double *bstats7) {
    Mat3 m = mat_in(H);
    put_bs(refine_homography(v2(x1, n), v2(x2, n), &m, cvt(opt)), bstats7);
    mat_out(m, H);
}
// ---- RANSAC drivers ---------------------------------------------------------------------------
void plo_ransac_pnp(const double *x, const double *X, uint64_t n, const plo_ransac_opt *opt, double max_error,
                    double *pose, char *inliers, plo_ransac_stats *stats, plo_counters *cnt) {
    CameraPose p = pose_in(pose);
    std::vector<char> m;
    Counters c;
    put(ransac_pnp(v2(x, n), v3(X, n), cvt(opt), max_error, &p, &m, &c), stats);
    pose_out(p, pose);
    mask_out(m, inliers);
    put(c, cnt);
}
void plo_ransac_relpose(const double *x1, const double *x2, uint64_t n, const plo_ransac_opt *opt, double max_error,
                        double *pose, char *inliers, plo_ransac_stats *stats, plo_counters *cnt) {
    CameraPose p = pose_in(pose);
    std::vector<char> m;
    Counters c;
    put(ransac_relpose(v2(x1, n), v2(x2, n), cvt(opt), max_error, &p, &m, &c), stats);
    pose_out(p, pose);
    mask_out(m, inliers);
    put(c, cnt);
}
void plo_ransac_fundamental(const double *x1, const double *x2, uint64_t n, const plo_ransac_opt *opt,
                            double max_error, int rfc, double *F, char *inliers, plo_ransac_stats *stats,
                            plo_counters *cnt) {
    Mat3 M = mat_in(F);
    std::vector<char> m;
    Counters c;
    put(ransac_fundamental(v2(x1, n), v2(x2, n), cvt(opt), max_error, rfc != 0, &M, &m, &c), stats);
    mat_out(M, F);
    mask_out(m, inliers);
    put(c, cnt);
}
void plo_ransac_homography(const double *x1, const double *x2, uint64_t n, const plo_ransac_opt *opt, double max_error,
                           double *H, char *inliers, plo_ransac_stats *stats, plo_counters *cnt) {
    Mat3 M = mat_in(H);
    std::vector<char> m;
    Counters c;
    put(ransac_homography(v2(x1, n), v2(x2, n), cvt(opt), max_error, &M, &m, &c), stats);
    mat_out(M, H);
    mask_out(m, inliers);
    put(c, cnt);
}
// ---- camera models (cam9 = model id followed by 8 parameter slots) ----------------------------
static Camera cam_in(const double *cam9) {
    Camera c;
    c.model_id = (int)cam9[0];
    for (int i = 0; i < 8; ++i) c.params[i] = cam9[1 + i];
    return c;
}
// out: n x (3 bearing + 6 M row-major 3x2)
void plo_camera_unproject_with_jac(const double *cam9, const double *xp, uint64_t n, double *out) {
    const Camera c = cam_in(cam9);
    for (uint64_t k = 0; k < n; ++k) {
        Vec2 p;
        p[0] = xp[2 * k];
        p[1] = xp[2 * k + 1];
        Vec3 d;
        double M[3][2];
        c.unproject_with_jac(p, &d, M);
        for (int i = 0; i < 3; ++i) out[9 * k + i] = d[i];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j) out[9 * k + 3 + 2 * i + j] = M[i][j];
    }
}
// out: n x 2 (Camera::unproject to 2D, camera_models.h:98-102)
void plo_camera_unproject2(const double *cam9, const double *xp, uint64_t n, double *out) {
    const Camera c = cam_in(cam9);
    for (uint64_t k = 0; k < n; ++k) {
        Vec2 p;
        p[0] = xp[2 * k];
        p[1] = xp[2 * k + 1];
        const Vec2 r = c.unproject2(p);
        out[2 * k] = r[0];
        out[2 * k + 1] = r[1];
    }
}
// out: n x (2 projection + 6 Jacobian row-major 2x3); proj_only: n x 2 from Camera::project
void plo_camera_project_with_jac(const double *cam9, const double *X, uint64_t n, double *out, double *proj_only) {
    const Camera c = cam_in(cam9);
    for (uint64_t k = 0; k < n; ++k) {
        const Vec3 x = mk3(X[3 * k], X[3 * k + 1], X[3 * k + 2]);
        Vec2 xp;
        double J[2][3];
        c.project_with_jac(x, &xp, J);
        out[8 * k] = xp[0];
        out[8 * k + 1] = xp[1];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 3; ++j) out[8 * k + 2 + 3 * i + j] = J[i][j];
        c.project(x, &xp);
        proj_only[2 * k] = xp[0];
        proj_only[2 * k + 1] = xp[1];
    }
}
double plo_camera_focal(const double *cam9) { return cam_in(cam9).focal(); }

// ---- tangent Sampson path (d: n x 3 bearings, M: n x 6 row-major 3x2) -------------------------
static std::vector<Mat32> m32(const double *M, uint64_t n) {
    std::vector<Mat32> r(n);
    for (uint64_t k = 0; k < n; ++k)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j) r[k].m[i][j] = M[6 * k + 2 * i + j];
    return r;
}
double plo_score_tangent(const double *pose, const double *d1, const double *d2, const double *M1, const double *M2,
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
void plo_refine_relpose_tangent(const double *d1, const double *d2, const double *M1, const double *M2, uint64_t n,
                                const plo_bundle_opt *bopt, double *pose, double *bstats7) {
    CameraPose p = pose_in(pose);
    const BundleStats st = refine_relpose(v3(d1, n), v3(d2, n), m32(M1, n), m32(M2, n), &p, cvt(bopt));
    pose_out(p, pose);
    put_bs(st, bstats7);
}
void plo_ransac_relpose_cameras(const double *x1, const double *x2, uint64_t n, const double *cam1_9,
                                const double *cam2_9, const plo_ransac_opt *opt, double max_error, double *pose,
                                char *inliers, plo_ransac_stats *stats, plo_counters *cnt) {
    CameraPose p = pose_in(pose);
    std::vector<char> m(n, 0);
    Counters c;
    put(ransac_relpose(v2(x1, n), v2(x2, n), cam_in(cam1_9), cam_in(cam2_9), cvt(opt), max_error, &p, &m, &c), stats);
    pose_out(p, pose);
    mask_out(m, inliers);
    put(c, cnt);
}
// ---- estimate_* -------------------------------------------------------------------------------
void plo_estimate_absolute_pose(const double *x, const double *X, uint64_t n, const plo_ransac_opt *ropt,
                                const plo_bundle_opt *bopt, double max_error, const double *cam9, double *pose,
                                char *inliers, plo_ransac_stats *stats, plo_counters *cnt) {
    Camera cam = cam_in(cam9);
    CameraPose p = pose_in(pose);
    std::vector<char> m(n, 0);
    Counters c;
    put(estimate_absolute_pose(v2(x, n), v3(X, n), cvt(ropt), cvt(bopt), max_error, cam, &p, &m, &c), stats);
    pose_out(p, pose);
    mask_out(m, inliers);
    put(c, cnt);
}
void plo_estimate_relative_pose(const double *x1, const double *x2, uint64_t n, const double *cam1_9,
                                const double *cam2_9, const plo_ransac_opt *ropt, const plo_bundle_opt *bopt,
                                double max_error, int tangent_sampson, double *pose, char *inliers,
                                plo_ransac_stats *stats, plo_counters *cnt) {
    Camera c1 = cam_in(cam1_9), c2 = cam_in(cam2_9);
    CameraPose p = pose_in(pose);
    std::vector<char> m(n, 0);
    Counters c;
    put(estimate_relative_pose(v2(x1, n), v2(x2, n), c1, c2, cvt(ropt), cvt(bopt), max_error, &p, &m, &c,
                               tangent_sampson != 0),
        stats);
    pose_out(p, pose);
    mask_out(m, inliers);
    put(c, cnt);
}
void plo_estimate_fundamental(const double *x1, const double *x2, uint64_t n, const plo_ransac_opt *ropt,
                              const plo_bundle_opt *bopt, double max_error, int rfc, double *F, char *inliers,
                              plo_ransac_stats *stats, plo_counters *cnt) {
    Mat3 M = mat_in(F);
    std::vector<char> m(n, 0);
    Counters c;
    put(estimate_fundamental(v2(x1, n), v2(x2, n), cvt(ropt), cvt(bopt), max_error, rfc != 0, &M, &m, &c), stats);
    mat_out(M, F);
    mask_out(m, inliers);
    put(c, cnt);
}
void plo_estimate_homography(const double *x1, const double *x2, uint64_t n, const plo_ransac_opt *ropt,
                             const plo_bundle_opt *bopt, double max_error, double *H, char *inliers,
                             plo_ransac_stats *stats, plo_counters *cnt) {
    Mat3 M = mat_in(H);
    std::vector<char> m(n, 0);
    Counters c;
    put(estimate_homography(v2(x1, n), v2(x2, n), cvt(ropt), cvt(bopt), max_error, &M, &m, &c), stats);
    mat_out(M, H);
    mask_out(m, inliers);
    put(c, cnt);
}

// ---- multi-threaded batch of relpose RANSAC problems (bench.py --impl reference / cpu_baseline) -----
// problems are packed back to back: x1[off[i]..off[i+1]) ; one problem per thread at a time.
double plo_ransac_relpose_batch_mt(const double *x1, const double *x2, const uint64_t *off, uint64_t count,
                                   const plo_ransac_opt *opts, const double *max_errors, int threads,
                                   double *poses, plo_ransac_stats *stats, plo_counters *cnts) {
    auto t0 = std::chrono::steady_clock::now();
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        for (;;) {
            uint64_t i = next.fetch_add(1);
            if (i >= count) break;
            const uint64_t n = off[i + 1] - off[i];
            plo_ransac_relpose(x1 + 2 * off[i], x2 + 2 * off[i], n, &opts[i], max_errors[i], poses + 7 * i, nullptr,
                               &stats[i], &cnts[i]);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, threads); ++t) th.emplace_back(work);
    for (auto &t : th) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

} // extern "C"
