// ORACLE — TEST INFRASTRUCTURE ONLY (see plo_math.h header).
// PARITY PARTLY PINNED: sampler, loop control flow, iteration arithmetic, univariate / p3p scalar solvers, Sturm root isolation, F / H scorers, masks, the real-focal check and the scalar camera code against the reference's own code (oracle/_ref, oracle/ref/ref_capi.cc); the transcription of PoseLib's logic for the WHOLE path (solvers, scorers, refiners, estimators, estimate_*) against the reference's own sources run on mini-Eigen (oracle/_ref/libplref2.so, oracle/ref/ref2_capi.cc, tests/test_ref_sources.py); Eigen's own arithmetic (reduction order, decompositions) is UNPINNED.
// CPU restatement of the PoseLib LO-RANSAC hot path; every function cites the reference file:line
// (relative to /root/reference) it follows.  Kept free of GPU-motivated changes on purpose.
#pragma once
#include "plo_math.h"
#include <vector>

namespace plo {
// test hook state (plo_set_reference_order, plo_solvers.cc): true while the reference's operation order is switched on
bool reference_order_enabled();

// ---- PoseLib/types.h:39-58 -------------------------------------------------------------------
struct RansacOptions {
    size_t max_iterations = 100000;
    size_t min_iterations = 1000;
    double dyn_num_trials_mult = 3.0;
    double success_prob = 0.9999;
    unsigned long seed = 0;
    bool progressive_sampling = false;
    size_t max_prosac_iterations = 100000;
    bool score_initial_model = false;
};
struct RansacStats {
    size_t refinements = 0;
    size_t iterations = 0;
    size_t num_inliers = 0;
    double inlier_ratio = 0;
    double model_score = std::numeric_limits<double>::max();
};
// PoseLib/types.h:60-106 (only the fields the four in-scope refiners read)
struct BundleOptions {
    size_t max_iterations = 100;
    enum LossType { TRIVIAL, TRUNCATED, HUBER, CAUCHY } loss_type = CAUCHY;
    double loss_scale = 1.0;
    double gradient_tol = 1e-12;
    double step_tol = 1e-8;
    double relative_cost_tol = 1e-10;
    double initial_lambda = 1e-3;
    double min_lambda = 1e-10;
    double max_lambda = 1e10;
};
struct BundleStats {
    size_t iterations = 0;
    double initial_cost = 0, cost = 0, lambda = 0, nu = 2.0;
    size_t invalid_steps = 0;
    double step_norm = 0, grad_norm = 0;
};

// extra counters for the metric: not part of the reference structs
struct Counters {
    size_t samples = 0, hypotheses = 0, scored_corrs = 0, lo_calls = 0;
    double lo_seconds = 0;
};

// ---- sampling (robust/sampling.{h,cc}) -------------------------------------------------------
int random_int(uint64_t &state);
void draw_sample(size_t sample_sz, size_t N, std::vector<size_t> *sample, uint64_t &rng);
struct RandomSampler {
    RandomSampler(size_t N, size_t K, const RansacOptions &opt);
    void generate_sample(std::vector<size_t> *sample);
    void initialize_prosac();
    size_t num_data, sample_sz;
    uint64_t state;
    bool use_prosac;
    size_t max_prosac_iterations, sample_k = 0, subset_sz = 0;
    std::vector<size_t> growth;
};

// ---- misc/univariate.cc, misc/sturm.h, misc/essential.cc -------------------------------------
bool solve_cubic_single_real(double c2, double c1, double c0, double &root);
int solve_cubic_real(double c2, double c1, double c0, double roots[3]);
int solve_quadratic_real(double a, double b, double c, double roots[2]);
int bisect_sturm10(const double *coeffs, double *roots, double tol = 1e-10);
void essential_from_motion(const CameraPose &pose, Mat3 *E);
bool check_cheirality(const CameraPose &pose, const Vec3 &x1, const Vec3 &x2, double min_depth = 0.0);
void motion_from_essential(const Mat3 &E, const std::vector<Vec3> &x1, const std::vector<Vec3> &x2,
                           std::vector<CameraPose> *poses);

// ---- solvers/ (inputs are unit bearing vectors) ----------------------------------------------
int p3p(const std::vector<Vec3> &x, const std::vector<Vec3> &X, std::vector<CameraPose> *out);
int p3p_lambdatwist(const std::vector<Vec3> &x, const std::vector<Vec3> &X, std::vector<CameraPose> *out); // p3p_lambdatwist.cc:68
int relpose_5pt(const std::vector<Vec3> &x1, const std::vector<Vec3> &x2, std::vector<Mat3> *E);
int relpose_5pt(const std::vector<Vec3> &x1, const std::vector<Vec3> &x2, std::vector<CameraPose> *out);
int relpose_7pt(const std::vector<Vec3> &x1, const std::vector<Vec3> &x2, std::vector<Mat3> *F);
int homography_4pt(const std::vector<Vec3> &x1, const std::vector<Vec3> &x2, Mat3 *H, bool check_cheirality = true);
// solvers/relpose_8pt.cc:52-95: non-minimal essential matrix from n >= 8 bearing pairs
void essential_matrix_8pt(const std::vector<Vec3> &x1, const std::vector<Vec3> &x2, Mat3 *E);
int relpose_8pt(const std::vector<Vec3> &x1, const std::vector<Vec3> &x2, std::vector<CameraPose> *output);

// ---- robust/utils.cc -------------------------------------------------------------------------
double compute_msac_score(const CameraPose &pose, const std::vector<Vec2> &x, const std::vector<Vec3> &X,
                          double sq_threshold, size_t *inlier_count);
double compute_sampson_msac_score(const CameraPose &pose, const std::vector<Vec2> &x1, const std::vector<Vec2> &x2,
                                  double sq_threshold, size_t *inlier_count);
double compute_sampson_msac_score(const Mat3 &F, const std::vector<Vec2> &x1, const std::vector<Vec2> &x2,
                                  double sq_threshold, size_t *inlier_count);
double compute_homography_msac_score(const Mat3 &H, const std::vector<Vec2> &x1, const std::vector<Vec2> &x2,
                                     double sq_threshold, size_t *inlier_count);
void get_inliers(const CameraPose &pose, const std::vector<Vec2> &x, const std::vector<Vec3> &X, double sq_threshold,
                 std::vector<char> *inliers);
int get_inliers(const CameraPose &pose, const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, double sq_threshold,
                std::vector<char> *inliers);
int get_inliers(const Mat3 &F, const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, double sq_threshold,
                std::vector<char> *inliers);
void get_homography_inliers(const Mat3 &H, const std::vector<Vec2> &x1, const std::vector<Vec2> &x2,
                            double sq_threshold, std::vector<char> *inliers);
struct Mat32 { double m[3][2]; }; // Eigen::Matrix<double,3,2> (unprojection Jacobian d bearing / d pixel)
double compute_tangent_sampson_msac_score(const CameraPose &pose, const std::vector<Vec3> &d1,
                                          const std::vector<Vec3> &d2, const std::vector<Mat32> &M1,
                                          const std::vector<Mat32> &M2, double sq_threshold, size_t *inlier_count);
int get_tangent_sampson_inliers(const CameraPose &pose, const std::vector<Vec3> &d1, const std::vector<Vec3> &d2,
                                const std::vector<Mat32> &M1, const std::vector<Mat32> &M2, double sq_threshold,
                                std::vector<char> *inliers);
double normalize_points(std::vector<Vec2> &x1, std::vector<Vec2> &x2, Mat3 &T1, Mat3 &T2, bool normalize_scale,
                        bool normalize_centroid, bool shared_scale);
bool calculate_RFC(const Mat3 &F);

// ---- robust/ransac_impl.h --------------------------------------------------------------------
double all_inlier_sample_probability(size_t num_inliers, size_t num_data, size_t sample_sz);
size_t compute_dynamic_max_iter(size_t num_inliers, size_t num_data, size_t sample_sz, double log_prob_missing_model,
                                double dyn_num_trials_mult, size_t min_iterations, size_t max_iterations);
// mock-estimator run used by the ransac_test.cc KATs (tests/ransac_test.cc:12-28)
RansacStats ransac_mock(size_t num_data, size_t sample_sz, size_t inlier_count, const RansacOptions &opt);

// ---- robust/bundle.cc entry points (LM refiners) ---------------------------------------------
// misc/camera_models.h:39-157 restricted to the six models on the path (ids = CameraModelId)
enum { CAM_NULL = -1, CAM_SIMPLE_PINHOLE = 0, CAM_PINHOLE = 1, CAM_SIMPLE_RADIAL = 2, CAM_RADIAL = 3, CAM_OPENCV = 4 };
struct Camera {
    int model_id = CAM_NULL;
    double params[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int num_params() const;
    double focal() const;
    void rescale(double scale);
    void project(const Vec3 &x, Vec2 *xp) const;
    void project_with_jac(const Vec3 &x, Vec2 *xp, double jac[2][3]) const;
    void unproject(const Vec2 &xp, Vec3 *x) const;
    Vec2 unproject2(const Vec2 &xp) const;
    void unproject_with_jac(const Vec2 &xp, Vec3 *x, double M[3][2]) const;
};
BundleStats bundle_adjust(const std::vector<Vec2> &x, const std::vector<Vec3> &X, CameraPose *pose,
                          const BundleOptions &opt);
BundleStats bundle_adjust_camera(const std::vector<Vec2> &x, const std::vector<Vec3> &X, const Camera &cam,
                                 CameraPose *pose, const BundleOptions &opt);
BundleStats refine_relpose(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, CameraPose *pose,
                           const BundleOptions &opt);
BundleStats refine_relpose(const std::vector<Vec3> &d1, const std::vector<Vec3> &d2, const std::vector<Mat32> &M1,
                           const std::vector<Mat32> &M2, CameraPose *pose, const BundleOptions &opt);
BundleStats refine_fundamental(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, Mat3 *F,
                               const BundleOptions &opt);
BundleStats refine_homography(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, Mat3 *H,
                              const BundleOptions &opt);

// ---- robust/ransac.cc drivers ----------------------------------------------------------------
RansacStats ransac_pnp(const std::vector<Vec2> &x, const std::vector<Vec3> &X, const RansacOptions &ropt,
                       double max_error, CameraPose *best, std::vector<char> *inliers, Counters *cnt = nullptr);
RansacStats ransac_relpose(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const RansacOptions &ropt,
                           double max_error, CameraPose *best, std::vector<char> *inliers, Counters *cnt = nullptr);
struct Camera;
RansacStats ransac_relpose(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const Camera &camera1,
                           const Camera &camera2, const RansacOptions &ropt, double max_error, CameraPose *best,
                           std::vector<char> *inliers, Counters *cnt = nullptr);
RansacStats ransac_fundamental(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const RansacOptions &ropt,
                               double max_error, bool real_focal_check, Mat3 *best, std::vector<char> *inliers,
                               Counters *cnt = nullptr);
RansacStats ransac_homography(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const RansacOptions &ropt,
                              double max_error, Mat3 *best, std::vector<char> *inliers, Counters *cnt = nullptr);

// ---- PoseLib/robust.cc entry points ----------------------------------------------------------
RansacStats estimate_absolute_pose(const std::vector<Vec2> &x, const std::vector<Vec3> &X, const RansacOptions &ropt,
                                   const BundleOptions &bopt, double max_error, const Camera &cam,
                                   CameraPose *pose, std::vector<char> *inliers, Counters *cnt = nullptr);
RansacStats estimate_relative_pose(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const Camera &cam1,
                                   const Camera &cam2, const RansacOptions &ropt, const BundleOptions &bopt,
                                   double max_error, CameraPose *pose, std::vector<char> *inliers,
                                   Counters *cnt = nullptr, bool tangent_sampson = false);
RansacStats estimate_fundamental(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const RansacOptions &ropt,
                                 const BundleOptions &bopt, double max_error, bool real_focal_check, Mat3 *F,
                                 std::vector<char> *inliers, Counters *cnt = nullptr);
RansacStats estimate_homography(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const RansacOptions &ropt,
                                const BundleOptions &bopt, double max_error, Mat3 *H, std::vector<char> *inliers,
                                Counters *cnt = nullptr);

} // namespace plo
