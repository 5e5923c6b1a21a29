// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY PARTLY PINNED: sampler, loop control flow, iteration arithmetic, univariate / p3p scalar solvers, Sturm root isolation, F / H scorers, masks, the real-focal check and the scalar camera code against the reference's own code (oracle/_ref, oracle/ref/ref_capi.cc); the transcription of PoseLib's logic for the WHOLE path (solvers, scorers, refiners, estimators, estimate_*) against the reference's own sources run on mini-Eigen (oracle/_ref/libplref2.so, oracle/ref/ref2_capi.cc, tests/test_ref_sources.py); Eigen's own arithmetic (reduction order, decompositions) is UNPINNED.
// Sampler, MSAC scorers, inlier masks, the generic LO-RANSAC loop and the four drivers,
// restated from PoseLib (paths relative to /root/reference).
#include "plo.h"
#include <chrono>

namespace plo {

// ============================ robust/sampling.cc ==============================================
// sampling.cc:37-43 — splitmix64 truncated to int
int random_int(uint64_t &state) {
    state += 0x9e3779b97f4a7c15ULL;
    uint64_t z = state;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return (int)(z ^ (z >> 31));
}
// sampling.cc:46-61 — note `int % size_t` sign-extends negatives before the modulo
void draw_sample(size_t sample_sz, size_t N, std::vector<size_t> *sample, uint64_t &rng) {
    for (size_t i = 0; i < sample_sz; ++i) {
        bool done = false;
        while (!done) {
            (*sample)[i] = random_int(rng) % N;
            done = true;
            for (size_t j = 0; j < i; ++j)
                if ((*sample)[i] == (*sample)[j]) {
                    done = false;
                    break;
                }
        }
    }
}
// sampling.h:51-58
RandomSampler::RandomSampler(size_t N, size_t K, const RansacOptions &opt)
    : num_data(N), sample_sz(K), state(opt.seed), use_prosac(opt.progressive_sampling),
      max_prosac_iterations(opt.max_prosac_iterations) {
    if (use_prosac) initialize_prosac();
}
// sampling.cc:85-103
void RandomSampler::generate_sample(std::vector<size_t> *sample) {
    if (use_prosac && sample_k < max_prosac_iterations) {
        draw_sample(sample_sz - 1, subset_sz - 1, sample, state);
        (*sample)[sample_sz - 1] = subset_sz - 1;
        sample_k++;
        if (sample_k < max_prosac_iterations) {
            if (sample_k > growth[subset_sz - 1]) {
                if (++subset_sz > num_data) subset_sz = num_data;
            }
        }
    } else {
        draw_sample(sample_sz, num_data, sample, state);
    }
}
// sampling.cc:105-136
void RandomSampler::initialize_prosac() {
    growth.assign(std::max(num_data, sample_sz), 0);
    double T_n = (double)max_prosac_iterations;
    for (size_t i = 0; i < sample_sz; ++i) T_n *= static_cast<double>(sample_sz - i) / (num_data - i);
    for (size_t n = 0; n < sample_sz; ++n) growth[n] = 1;
    size_t T_np = 1;
    for (size_t n = sample_sz; n < num_data; ++n) {
        const double T_n_next = T_n * (n + 1.0) / (n + 1.0 - sample_sz);
        T_np += std::ceil(T_n_next - T_n); // size_t += double  (converted through double)
        growth[n] = T_np;
        T_n = T_n_next;
    }
    sample_k = 1;
    subset_sz = sample_sz;
}

// ============================ robust/utils.cc scorers =========================================
// utils.cc:36-63
double compute_msac_score(const CameraPose &pose, const std::vector<Vec2> &x, const std::vector<Vec3> &X,
                          double sq_threshold, size_t *inlier_count) {
    *inlier_count = 0;
    double score = 0.0;
    const Mat3 R = pose.R();
    const double P0_0 = R(0, 0), P0_1 = R(0, 1), P0_2 = R(0, 2), P0_3 = pose.t[0];
    const double P1_0 = R(1, 0), P1_1 = R(1, 1), P1_2 = R(1, 2), P1_3 = pose.t[1];
    const double P2_0 = R(2, 0), P2_1 = R(2, 1), P2_2 = R(2, 2), P2_3 = pose.t[2];
    for (size_t k = 0; k < x.size(); ++k) {
        const double X0 = X[k][0], X1 = X[k][1], X2 = X[k][2];
        const double x0 = x[k][0], x1 = x[k][1];
        const double z0 = P0_0 * X0 + P0_1 * X1 + P0_2 * X2 + P0_3;
        const double z1 = P1_0 * X0 + P1_1 * X1 + P1_2 * X2 + P1_3;
        const double z2 = P2_0 * X0 + P2_1 * X1 + P2_2 * X2 + P2_3;
        if (z2 <= 0.0) continue;
        const double inv_z2 = 1.0 / z2;
        const double r_0 = z0 * inv_z2 - x0;
        const double r_1 = z1 * inv_z2 - x1;
        const double r_sq = r_0 * r_0 + r_1 * r_1;
        if (r_sq < sq_threshold) {
            (*inlier_count)++;
            score += r_sq;
        }
    }
    score += (x.size() - *inlier_count) * sq_threshold;
    return score;
}

namespace {
struct SampsonTerms {
    double r2;
};
// shared residual of utils.cc:171-185 / :216-230 / :446-461 / :491-505
inline double sampson_r2(const Mat3 &E, const Vec2 &p1, const Vec2 &p2) {
    const double x1_0 = p1[0], x1_1 = p1[1], x2_0 = p2[0], x2_1 = p2[1];
    const double Ex1_0 = E(0, 0) * x1_0 + E(0, 1) * x1_1 + E(0, 2);
    const double Ex1_1 = E(1, 0) * x1_0 + E(1, 1) * x1_1 + E(1, 2);
    const double Ex1_2 = E(2, 0) * x1_0 + E(2, 1) * x1_1 + E(2, 2);
    const double Ex2_0 = E(0, 0) * x2_0 + E(1, 0) * x2_1 + E(2, 0);
    const double Ex2_1 = E(0, 1) * x2_0 + E(1, 1) * x2_1 + E(2, 1);
    const double C = x2_0 * Ex1_0 + x2_1 * Ex1_1 + Ex1_2;
    const double Cx = Ex1_0 * Ex1_0 + Ex1_1 * Ex1_1;
    const double Cy = Ex2_0 * Ex2_0 + Ex2_1 * Ex2_1;
    return C * C / (Cx + Cy);
}
} // namespace

// utils.cc:158-201
double compute_sampson_msac_score(const CameraPose &pose, const std::vector<Vec2> &x1, const std::vector<Vec2> &x2,
                                  double sq_threshold, size_t *inlier_count) {
    *inlier_count = 0;
    Mat3 E;
    essential_from_motion(pose, &E);
    double score = 0.0;
    for (size_t k = 0; k < x1.size(); ++k) {
        const double r2 = sampson_r2(E, x1[k], x2[k]);
        if (r2 < sq_threshold) {
            const bool cheirality = check_cheirality(pose, bearing(x1[k]), bearing(x2[k]), 0.01);
            if (cheirality) {
                (*inlier_count)++;
                score += r2;
            } else {
                score += sq_threshold;
            }
        } else {
            score += sq_threshold;
        }
    }
    return score;
}
namespace {
// r2 of utils.cc:282-285 / :554-557; `M^T * E * d` is evaluated by Eigen as (M^T E) d
inline double tangent_sampson_r2(const Mat3 &E, const Vec3 &d1, const Vec3 &d2, const Mat32 &M1, const Mat32 &M2) {
    const Vec3 Ed1 = E * d1;
    const double C = d2[0] * Ed1[0] + d2[1] * Ed1[1] + d2[2] * Ed1[2];
    double a[2], b[2];
    for (int i = 0; i < 2; ++i) {
        double T1[3], T2[3];
        for (int j = 0; j < 3; ++j) {
            T2[j] = M2.m[0][i] * E(0, j) + M2.m[1][i] * E(1, j) + M2.m[2][i] * E(2, j);
            T1[j] = M1.m[0][i] * E(j, 0) + M1.m[1][i] * E(j, 1) + M1.m[2][i] * E(j, 2);
        }
        a[i] = T2[0] * d1[0] + T2[1] * d1[1] + T2[2] * d1[2];
        b[i] = T1[0] * d2[0] + T1[1] * d2[1] + T1[2] * d2[2];
    }
    const double denom2 = (a[0] * a[0] + a[1] * a[1]) + (b[0] * b[0] + b[1] * b[1]);
    return C * C / denom2;
}
} // namespace
// utils.cc:269-298
double compute_tangent_sampson_msac_score(const CameraPose &pose, const std::vector<Vec3> &d1,
                                          const std::vector<Vec3> &d2, const std::vector<Mat32> &M1,
                                          const std::vector<Mat32> &M2, double sq_threshold, size_t *inlier_count) {
    Mat3 E;
    essential_from_motion(pose, &E);
    *inlier_count = 0;
    double score = 0;
    for (size_t i = 0; i < d1.size(); ++i) {
        const double r2 = tangent_sampson_r2(E, d1[i], d2[i], M1[i], M2[i]);
        if (r2 < sq_threshold) {
            const bool cheirality = check_cheirality(pose, d1[i], d2[i], 0.01);
            if (cheirality) {
                (*inlier_count)++;
                score += r2;
            } else {
                score += sq_threshold;
            }
        } else {
            score += sq_threshold;
        }
    }
    return score;
}
// utils.cc:541-569
int get_tangent_sampson_inliers(const CameraPose &pose, const std::vector<Vec3> &d1, const std::vector<Vec3> &d2,
                                const std::vector<Mat32> &M1, const std::vector<Mat32> &M2, double sq_threshold,
                                std::vector<char> *inliers) {
    Mat3 E;
    essential_from_motion(pose, &E);
    inliers->resize(d1.size());
    size_t inlier_count = 0;
    for (size_t i = 0; i < d1.size(); ++i) {
        const double r2 = tangent_sampson_r2(E, d1[i], d2[i], M1[i], M2[i]);
        bool inlier = (r2 < sq_threshold);
        if (inlier) {
            const bool cheirality = check_cheirality(pose, d1[i], d2[i], 0.01);
            if (cheirality) inlier_count++;
            else inlier = false;
        }
        (*inliers)[i] = inlier;
    }
    return (int)inlier_count;
}
// utils.cc:204-239
double compute_sampson_msac_score(const Mat3 &F, const std::vector<Vec2> &x1, const std::vector<Vec2> &x2,
                                  double sq_threshold, size_t *inlier_count) {
    *inlier_count = 0;
    double score = 0.0;
    for (size_t k = 0; k < x1.size(); ++k) {
        const double r2 = sampson_r2(F, x1[k], x2[k]);
        if (r2 < sq_threshold) {
            (*inlier_count)++;
            score += r2;
        } else {
            score += sq_threshold;
        }
    }
    return score;
}
namespace {
inline double homography_r2(const Mat3 &H, const Vec2 &p1, const Vec2 &p2) { // utils.cc:310-320
    const double Hx1_0 = H(0, 0) * p1[0] + H(0, 1) * p1[1] + H(0, 2);
    const double Hx1_1 = H(1, 0) * p1[0] + H(1, 1) * p1[1] + H(1, 2);
    const double inv_Hx1_2 = 1.0 / (H(2, 0) * p1[0] + H(2, 1) * p1[1] + H(2, 2));
    const double r0 = Hx1_0 * inv_Hx1_2 - p2[0];
    const double r1 = Hx1_1 * inv_Hx1_2 - p2[1];
    return r0 * r0 + r1 * r1;
}
} // namespace
// utils.cc:300-329
double compute_homography_msac_score(const Mat3 &H, const std::vector<Vec2> &x1, const std::vector<Vec2> &x2,
                                     double sq_threshold, size_t *inlier_count) {
    *inlier_count = 0;
    double score = 0;
    for (size_t k = 0; k < x1.size(); ++k) {
        const double r2 = homography_r2(H, x1[k], x2[k]);
        if (r2 < sq_threshold) {
            (*inlier_count)++;
            score += r2;
        } else {
            score += sq_threshold;
        }
    }
    return score;
}
// utils.cc:331-351
void get_homography_inliers(const Mat3 &H, const std::vector<Vec2> &x1, const std::vector<Vec2> &x2,
                            double sq_threshold, std::vector<char> *inliers) {
    inliers->resize(x1.size());
    for (size_t k = 0; k < x1.size(); ++k) (*inliers)[k] = (homography_r2(H, x1[k], x2[k]) < sq_threshold);
}
// utils.cc:374-383 — Eigen expression order: Z = R*X + t ; hnormalized ; squaredNorm
void get_inliers(const CameraPose &pose, const std::vector<Vec2> &x, const std::vector<Vec3> &X, double sq_threshold,
                 std::vector<char> *inliers) {
    inliers->resize(x.size());
    const Mat3 R = pose.R();
    for (size_t k = 0; k < x.size(); ++k) {
        const Vec3 Z = R * X[k] + pose.t;
        const double d0 = Z[0] / Z[2] - x[k][0], d1 = Z[1] / Z[2] - x[k][1];
        const double r2 = d0 * d0 + d1 * d1;
        (*inliers)[k] = (r2 < sq_threshold && Z[2] > 0.0);
    }
}
// utils.cc:434-476
int get_inliers(const CameraPose &pose, const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, double sq_threshold,
                std::vector<char> *inliers) {
    inliers->resize(x1.size());
    Mat3 E;
    essential_from_motion(pose, &E);
    size_t inlier_count = 0;
    for (size_t k = 0; k < x1.size(); ++k) {
        const double r2 = sampson_r2(E, x1[k], x2[k]);
        bool inlier = (r2 < sq_threshold);
        if (inlier) {
            if (check_cheirality(pose, bearing(x1[k]), bearing(x2[k]), 0.01)) inlier_count++;
            else inlier = false;
        }
        (*inliers)[k] = inlier;
    }
    return (int)inlier_count;
}
// utils.cc:479-513
int get_inliers(const Mat3 &F, const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, double sq_threshold,
                std::vector<char> *inliers) {
    inliers->resize(x1.size());
    size_t inlier_count = 0;
    for (size_t k = 0; k < x1.size(); ++k) {
        const bool inlier = (sampson_r2(F, x1[k], x2[k]) < sq_threshold);
        if (inlier) inlier_count++;
        (*inliers)[k] = inlier;
    }
    return (int)inlier_count;
}
// utils.cc:584-644
double normalize_points(std::vector<Vec2> &x1, std::vector<Vec2> &x2, Mat3 &T1, Mat3 &T2, bool normalize_scale,
                        bool normalize_centroid, bool shared_scale) {
    T1 = mat3_identity();
    T2 = mat3_identity();
    const size_t n = x1.size();
    if (normalize_centroid) {
        double c1[2] = {0, 0}, c2[2] = {0, 0};
        for (size_t k = 0; k < n; ++k) {
            c1[0] += x1[k][0]; c1[1] += x1[k][1];
            c2[0] += x2[k][0]; c2[1] += x2[k][1];
        }
        c1[0] /= n; c1[1] /= n; c2[0] /= n; c2[1] /= n;
        T1(0, 2) = -c1[0]; T1(1, 2) = -c1[1];
        T2(0, 2) = -c2[0]; T2(1, 2) = -c2[1];
        for (size_t k = 0; k < n; ++k) {
            x1[k][0] -= c1[0]; x1[k][1] -= c1[1];
            x2[k][0] -= c2[0]; x2[k][1] -= c2[1];
        }
    }
    auto nrm = [](const Vec2 &p) { return std::sqrt(p[0] * p[0] + p[1] * p[1]); };
    if (normalize_scale && shared_scale) {
        double scale = 0.0;
        for (size_t k = 0; k < n; ++k) {
            scale += nrm(x1[k]);
            scale += nrm(x2[k]);
        }
        scale /= std::sqrt(2) * n;
        for (size_t k = 0; k < n; ++k) {
            x1[k][0] /= scale; x1[k][1] /= scale;
            x2[k][0] /= scale; x2[k][1] /= scale;
        }
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 3; ++c) {
                T1(r, c) *= 1.0 / scale;
                T2(r, c) *= 1.0 / scale;
            }
        return scale;
    } else if (normalize_scale && !shared_scale) {
        double scale1 = 0.0, scale2 = 0.0;
        for (size_t k = 0; k < n; ++k) {
            scale1 += nrm(x1[k]);
            scale2 += nrm(x2[k]);
        }
        scale1 /= n / std::sqrt(2);
        scale2 /= n / std::sqrt(2);
        for (size_t k = 0; k < n; ++k) {
            x1[k][0] /= scale1; x1[k][1] /= scale1;
            x2[k][0] /= scale2; x2[k][1] /= scale2;
        }
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 3; ++c) {
                T1(r, c) *= 1.0 / scale1;
                T2(r, c) *= 1.0 / scale2;
            }
        return std::sqrt(scale1 * scale2);
    }
    return 1.0;
}
// utils.cc:646-672 — NB: den/num are `float`
bool calculate_RFC(const Mat3 &F) {
    float den, num;
    den = F(0, 0) * F(0, 1) * F(2, 0) * F(2, 2) - F(0, 0) * F(0, 2) * F(2, 0) * F(2, 1) +
          F(0, 1) * F(0, 1) * F(2, 1) * F(2, 2) - F(0, 1) * F(0, 2) * F(2, 1) * F(2, 1) +
          F(1, 0) * F(1, 1) * F(2, 0) * F(2, 2) - F(1, 0) * F(1, 2) * F(2, 0) * F(2, 1) +
          F(1, 1) * F(1, 1) * F(2, 1) * F(2, 2) - F(1, 1) * F(1, 2) * F(2, 1) * F(2, 1);
    num = -F(2, 2) * (F(0, 1) * F(0, 2) * F(2, 2) - F(0, 2) * F(0, 2) * F(2, 1) + F(1, 1) * F(1, 2) * F(2, 2) -
                      F(1, 2) * F(1, 2) * F(2, 1));
    if (num * den < 0) return false;
    den = F(0, 0) * F(1, 0) * F(0, 2) * F(2, 2) - F(0, 0) * F(2, 0) * F(0, 2) * F(1, 2) +
          F(1, 0) * F(1, 0) * F(1, 2) * F(2, 2) - F(1, 0) * F(2, 0) * F(1, 2) * F(1, 2) +
          F(0, 1) * F(1, 1) * F(0, 2) * F(2, 2) - F(0, 1) * F(2, 1) * F(0, 2) * F(1, 2) +
          F(1, 1) * F(1, 1) * F(1, 2) * F(2, 2) - F(1, 1) * F(2, 1) * F(1, 2) * F(1, 2);
    num = -F(2, 2) * (F(1, 0) * F(2, 0) * F(2, 2) - F(2, 0) * F(2, 0) * F(1, 2) + F(1, 1) * F(2, 1) * F(2, 2) -
                      F(2, 1) * F(2, 1) * F(1, 2));
    if (num * den < 0) return false;
    return true;
}

// ============================ robust/ransac_impl.h ============================================
// ransac_impl.h:43-56
double all_inlier_sample_probability(size_t num_inliers, size_t num_data, size_t sample_sz) {
    if (sample_sz == 0) return 1.0;
    if (num_inliers < sample_sz || num_data < sample_sz) return 0.0;
    double p = 1.0;
    for (size_t i = 0; i < sample_sz; ++i) p *= static_cast<double>(num_inliers - i) / static_cast<double>(num_data - i);
    return p;
}
// ransac_impl.h:58-74
size_t compute_dynamic_max_iter(size_t num_inliers, size_t num_data, size_t sample_sz, double log_prob_missing_model,
                                double dyn_num_trials_mult, size_t min_iterations, size_t max_iterations) {
    const double p = all_inlier_sample_probability(num_inliers, num_data, sample_sz);
    if (p >= 0.9999) return min_iterations;
    if (p <= 0.0001) return max_iterations;
    const double prob_outlier = 1.0 - p;
    const size_t num_iters =
        static_cast<size_t>(std::ceil(log_prob_missing_model / std::log(prob_outlier) * dyn_num_trials_mult));
    return std::max(min_iterations, std::min(max_iterations, num_iters));
}

namespace {
struct RansacState { // ransac_impl.h:99-104
    size_t best_minimal_inlier_count = 0;
    double best_minimal_msac_score = std::numeric_limits<double>::max();
    size_t dynamic_max_iter = 100000;
    double log_prob_missing_model = std::log(1.0 - 0.9999);
};

// ransac_impl.h:106-154
template <typename Solver, typename Model>
void score_models(const Solver &estimator, const std::vector<Model> &models, const RansacOptions &opt,
                  RansacState &state, RansacStats &stats, Model *best_model) {
    int best_model_ind = -1;
    size_t inlier_count = 0;
    for (size_t i = 0; i < models.size(); ++i) {
        const double score_msac = estimator.score_model(models[i], &inlier_count);
        const bool more_inliers = inlier_count > state.best_minimal_inlier_count;
        const bool better_score = score_msac < state.best_minimal_msac_score;
        if (more_inliers || better_score) {
            if (more_inliers) state.best_minimal_inlier_count = inlier_count;
            if (better_score) state.best_minimal_msac_score = score_msac;
            best_model_ind = (int)i;
            if (score_msac < stats.model_score) {
                stats.model_score = score_msac;
                *best_model = models[i];
                stats.num_inliers = inlier_count;
            }
        }
    }
    if (best_model_ind == -1) return;
    Model refined_model = models[best_model_ind];
    estimator.refine_model(&refined_model);
    stats.refinements++;
    const double refined_msac_score = estimator.score_model(refined_model, &inlier_count);
    if (refined_msac_score < stats.model_score) {
        stats.model_score = refined_msac_score;
        stats.num_inliers = inlier_count;
        *best_model = refined_model;
    }
    stats.inlier_ratio = static_cast<double>(stats.num_inliers) / static_cast<double>(estimator.num_data);
    state.dynamic_max_iter =
        compute_dynamic_max_iter(stats.num_inliers, estimator.num_data, estimator.sample_sz,
                                 state.log_prob_missing_model, opt.dyn_num_trials_mult, opt.min_iterations,
                                 opt.max_iterations);
}

// ransac_impl.h:157-201
template <typename Solver, typename Model> RansacStats ransac(Solver &estimator, const RansacOptions &opt, Model *best_model) {
    RansacStats stats;
    if (estimator.num_data < estimator.sample_sz) return stats;
    stats.num_inliers = 0;
    stats.model_score = std::numeric_limits<double>::max();
    RansacState state;
    state.dynamic_max_iter = opt.max_iterations;
    state.log_prob_missing_model = std::log(1.0 - opt.success_prob);
    if (opt.score_initial_model) {
        std::vector<Model> init = {*best_model};
        score_models(estimator, init, opt, state, stats, best_model);
    }
    size_t inlier_count = 0;
    std::vector<Model> models;
    for (stats.iterations = 0; stats.iterations < opt.max_iterations; stats.iterations++) {
        if (stats.iterations > opt.min_iterations && stats.iterations > state.dynamic_max_iter) break;
        models.clear();
        estimator.generate_models(&models);
        score_models(estimator, models, opt, state, stats, best_model);
    }
    Model refined_model = *best_model;
    estimator.refine_model(&refined_model);
    stats.refinements++;
    const double refined_msac_score = estimator.score_model(refined_model, &inlier_count);
    if (refined_msac_score < stats.model_score) { // NB: model_score is NOT updated here (:195-198)
        *best_model = refined_model;
        stats.num_inliers = inlier_count;
    }
    return stats;
}

// tests/ransac_test.cc:12-28
struct MockEstimator {
    size_t sample_sz, num_data, inlier_count;
    void generate_models(std::vector<int> *models) const { models->push_back(0); }
    double score_model(const int &, size_t *c) const {
        *c = inlier_count;
        return 0.0;
    }
    void refine_model(int *) const {}
};

inline BundleOptions lo_bundle_options(double max_error) { // estimators/*.cc refine_model
    BundleOptions b;
    b.loss_type = BundleOptions::TRUNCATED;
    b.loss_scale = max_error;
    b.max_iterations = 25;
    return b;
}
struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double sec() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

// estimators/absolute_pose.{h,cc}:39-69
struct AbsolutePoseEstimator {
    AbsolutePoseEstimator(const RansacOptions &ropt, double max_err, const std::vector<Vec2> &x_,
                          const std::vector<Vec3> &X_, Counters *c)
        : sample_sz(3), num_data(x_.size()), max_error(max_err), x(x_), X(X_), sampler(num_data, sample_sz, ropt),
          cnt(c) {
        xs.resize(sample_sz);
        Xs.resize(sample_sz);
        sample.resize(sample_sz);
    }
    void generate_models(std::vector<CameraPose> *models) {
        models->clear();
        sampler.generate_sample(&sample);
        for (size_t k = 0; k < sample_sz; ++k) {
            xs[k] = bearing(x[sample[k]]);
            Xs[k] = X[sample[k]];
        }
        p3p(xs, Xs, models);
        if (cnt) cnt->samples++;
    }
    double score_model(const CameraPose &pose, size_t *inlier_count) const {
        if (cnt) { cnt->hypotheses++; cnt->scored_corrs += num_data; }
        return compute_msac_score(pose, x, X, max_error * max_error, inlier_count);
    }
    void refine_model(CameraPose *pose) const {
        Timer tm;
        bundle_adjust(x, X, pose, lo_bundle_options(max_error));
        if (cnt) { cnt->lo_calls++; cnt->lo_seconds += tm.sec(); }
    }
    size_t sample_sz, num_data;
    double max_error;
    const std::vector<Vec2> &x;
    const std::vector<Vec3> &X;
    RandomSampler sampler;
    std::vector<Vec3> xs, Xs;
    std::vector<size_t> sample;
    Counters *cnt;
};
// estimators/relative_pose.{h,cc}:40-86
struct RelativePoseEstimator {
    RelativePoseEstimator(const RansacOptions &ropt, double max_err, const std::vector<Vec2> &a,
                          const std::vector<Vec2> &b, Counters *c)
        : sample_sz(5), num_data(a.size()), max_error(max_err), x1(a), x2(b), sampler(num_data, sample_sz, ropt),
          cnt(c) {
        x1s.resize(sample_sz);
        x2s.resize(sample_sz);
        sample.resize(sample_sz);
    }
    void generate_models(std::vector<CameraPose> *models) {
        models->clear();
        sampler.generate_sample(&sample);
        for (size_t k = 0; k < sample_sz; ++k) {
            x1s[k] = bearing(x1[sample[k]]);
            x2s[k] = bearing(x2[sample[k]]);
        }
        relpose_5pt(x1s, x2s, models);
        if (cnt) cnt->samples++;
    }
    double score_model(const CameraPose &pose, size_t *inlier_count) const {
        if (cnt) { cnt->hypotheses++; cnt->scored_corrs += num_data; }
        return compute_sampson_msac_score(pose, x1, x2, max_error * max_error, inlier_count);
    }
    void refine_model(CameraPose *pose) const {
        Timer tm;
        std::vector<char> inliers;
        const int num_inl = get_inliers(*pose, x1, x2, 5 * (max_error * max_error), &inliers);
        if (num_inl > 5) {
            std::vector<Vec2> a, b;
            a.reserve(num_inl);
            b.reserve(num_inl);
            for (size_t k = 0; k < x1.size(); ++k)
                if (inliers[k]) {
                    a.push_back(x1[k]);
                    b.push_back(x2[k]);
                }
            refine_relpose(a, b, pose, lo_bundle_options(max_error));
        }
        if (cnt) { cnt->lo_calls++; cnt->lo_seconds += tm.sec(); }
    }
    size_t sample_sz, num_data;
    double max_error;
    const std::vector<Vec2> &x1, &x2;
    RandomSampler sampler;
    std::vector<Vec3> x1s, x2s;
    std::vector<size_t> sample;
    Counters *cnt;
};
// estimators/relative_pose.h:71-106, relative_pose.cc:88-108
struct CameraRelativePoseEstimator {
    CameraRelativePoseEstimator(const RansacOptions &ropt, double max_err, const std::vector<Vec2> &a,
                                const std::vector<Vec2> &b, const Camera &cam1, const Camera &cam2, Counters *c)
        : sample_sz(5), num_data(a.size()), max_error(max_err), sampler(num_data, sample_sz, ropt), cnt(c) {
        x1s.resize(sample_sz);
        x2s.resize(sample_sz);
        sample.resize(sample_sz);
        d1.resize(num_data);
        d2.resize(num_data);
        M1.resize(num_data);
        M2.resize(num_data);
        for (size_t k = 0; k < num_data; ++k) { // Camera::unproject_with_jac vector wrapper (camera_models.cc:271-296)
            cam1.unproject_with_jac(a[k], &d1[k], M1[k].m);
            cam2.unproject_with_jac(b[k], &d2[k], M2[k].m);
        }
    }
    void generate_models(std::vector<CameraPose> *models) {
        models->clear();
        sampler.generate_sample(&sample);
        for (size_t k = 0; k < sample_sz; ++k) {
            x1s[k] = d1[sample[k]];
            x2s[k] = d2[sample[k]];
        }
        relpose_5pt(x1s, x2s, models);
        if (cnt) cnt->samples++;
    }
    double score_model(const CameraPose &pose, size_t *inlier_count) const {
        if (cnt) { cnt->hypotheses++; cnt->scored_corrs += num_data; }
        return compute_tangent_sampson_msac_score(pose, d1, d2, M1, M2, max_error * max_error, inlier_count);
    }
    void refine_model(CameraPose *pose) const {
        Timer tm;
        refine_relpose(d1, d2, M1, M2, pose, lo_bundle_options(max_error));
        if (cnt) { cnt->lo_calls++; cnt->lo_seconds += tm.sec(); }
    }
    size_t sample_sz, num_data;
    double max_error;
    RandomSampler sampler;
    std::vector<Vec3> x1s, x2s, d1, d2;
    std::vector<Mat32> M1, M2;
    std::vector<size_t> sample;
    Counters *cnt;
};
// estimators/relative_pose.{h,cc}:309-336 / :384-412
struct FundamentalEstimator {
    FundamentalEstimator(const RansacOptions &ropt, double max_err, bool rfc, const std::vector<Vec2> &a,
                         const std::vector<Vec2> &b, Counters *c)
        : sample_sz(7), num_data(a.size()), max_error(max_err), real_focal_check(rfc), x1(a), x2(b),
          sampler(num_data, sample_sz, ropt), cnt(c) {
        x1s.resize(sample_sz);
        x2s.resize(sample_sz);
        sample.resize(sample_sz);
    }
    void generate_models(std::vector<Mat3> *models) {
        models->clear();
        sampler.generate_sample(&sample);
        for (size_t k = 0; k < sample_sz; ++k) {
            x1s[k] = bearing(x1[sample[k]]);
            x2s[k] = bearing(x2[sample[k]]);
        }
        relpose_7pt(x1s, x2s, models);
        if (real_focal_check) {
            for (int i = (int)models->size() - 1; i >= 0; i--)
                if (!calculate_RFC((*models)[i])) models->erase(models->begin() + i);
        }
        if (cnt) cnt->samples++;
    }
    double score_model(const Mat3 &F, size_t *inlier_count) const {
        if (cnt) { cnt->hypotheses++; cnt->scored_corrs += num_data; }
        return compute_sampson_msac_score(F, x1, x2, max_error * max_error, inlier_count);
    }
    void refine_model(Mat3 *F) const {
        Timer tm;
        refine_fundamental(x1, x2, F, lo_bundle_options(max_error));
        if (cnt) { cnt->lo_calls++; cnt->lo_seconds += tm.sec(); }
    }
    size_t sample_sz, num_data;
    double max_error;
    bool real_focal_check;
    const std::vector<Vec2> &x1, &x2;
    RandomSampler sampler;
    std::vector<Vec3> x1s, x2s;
    std::vector<size_t> sample;
    Counters *cnt;
};
// estimators/homography.{h,cc}:38-64 / :36-61
struct HomographyEstimator {
    HomographyEstimator(const RansacOptions &ropt, double max_err, const std::vector<Vec2> &a,
                        const std::vector<Vec2> &b, Counters *c)
        : sample_sz(4), num_data(a.size()), max_error(max_err), x1(a), x2(b), sampler(num_data, sample_sz, ropt),
          cnt(c) {
        x1s.resize(sample_sz);
        x2s.resize(sample_sz);
        sample.resize(sample_sz);
    }
    void generate_models(std::vector<Mat3> *models) {
        models->clear();
        sampler.generate_sample(&sample);
        for (size_t k = 0; k < sample_sz; ++k) {
            x1s[k] = bearing(x1[sample[k]]);
            x2s[k] = bearing(x2[sample[k]]);
        }
        Mat3 H;
        if (homography_4pt(x1s, x2s, &H, true) > 0) models->push_back(H);
        if (cnt) cnt->samples++;
    }
    double score_model(const Mat3 &H, size_t *inlier_count) const {
        if (cnt) { cnt->hypotheses++; cnt->scored_corrs += num_data; }
        return compute_homography_msac_score(H, x1, x2, max_error * max_error, inlier_count);
    }
    void refine_model(Mat3 *H) const {
        Timer tm;
        refine_homography(x1, x2, H, lo_bundle_options(max_error));
        if (cnt) { cnt->lo_calls++; cnt->lo_seconds += tm.sec(); }
    }
    size_t sample_sz, num_data;
    double max_error;
    const std::vector<Vec2> &x1, &x2;
    RandomSampler sampler;
    std::vector<Vec3> x1s, x2s;
    std::vector<size_t> sample;
    Counters *cnt;
};
} // namespace

RansacStats ransac_mock(size_t num_data, size_t sample_sz, size_t inlier_count, const RansacOptions &opt) {
    MockEstimator est{sample_sz, num_data, inlier_count};
    int best = -1;
    return ransac<MockEstimator, int>(est, opt, &best);
}

// ============================ robust/ransac.cc drivers ========================================
// ransac.cc:44-57
RansacStats ransac_pnp(const std::vector<Vec2> &x, const std::vector<Vec3> &X, const RansacOptions &ropt,
                       double max_error, CameraPose *best, std::vector<char> *inliers, Counters *cnt) {
    if (!ropt.score_initial_model) *best = CameraPose();
    AbsolutePoseEstimator est(ropt, max_error, x, X, cnt);
    RansacStats stats = ransac<AbsolutePoseEstimator, CameraPose>(est, ropt, best);
    get_inliers(*best, x, X, max_error * max_error, inliers);
    return stats;
}
// ransac.cc:142-154
RansacStats ransac_relpose(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const RansacOptions &ropt,
                           double max_error, CameraPose *best, std::vector<char> *inliers, Counters *cnt) {
    if (!ropt.score_initial_model) *best = CameraPose();
    RelativePoseEstimator est(ropt, max_error, x1, x2, cnt);
    RansacStats stats = ransac<RelativePoseEstimator, CameraPose>(est, ropt, best);
    get_inliers(*best, x1, x2, max_error * max_error, inliers);
    return stats;
}
// ransac.cc:155-168 (points in the pixel units of the two cameras; tangent Sampson error)
RansacStats ransac_relpose(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const Camera &camera1,
                           const Camera &camera2, const RansacOptions &ropt, double max_error, CameraPose *best,
                           std::vector<char> *inliers, Counters *cnt) {
    *best = CameraPose(); // :159-160 resets unconditionally
    CameraRelativePoseEstimator est(ropt, max_error, x1, x2, camera1, camera2, cnt);
    RansacStats stats = ransac<CameraRelativePoseEstimator, CameraPose>(est, ropt, best);
    get_tangent_sampson_inliers(*best, est.d1, est.d2, est.M1, est.M2, max_error * max_error, inliers);
    return stats;
}
// ransac.cc:248-262
RansacStats ransac_fundamental(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const RansacOptions &ropt,
                               double max_error, bool real_focal_check, Mat3 *best, std::vector<char> *inliers,
                               Counters *cnt) {
    if (!ropt.score_initial_model) *best = mat3_identity();
    FundamentalEstimator est(ropt, max_error, real_focal_check, x1, x2, cnt);
    RansacStats stats = ransac<FundamentalEstimator, Mat3>(est, ropt, best);
    get_inliers(*best, x1, x2, max_error * max_error, inliers);
    return stats;
}
// ransac.cc:300-314
RansacStats ransac_homography(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const RansacOptions &ropt,
                              double max_error, Mat3 *best, std::vector<char> *inliers, Counters *cnt) {
    if (!ropt.score_initial_model) *best = mat3_identity();
    HomographyEstimator est(ropt, max_error, x1, x2, cnt);
    RansacStats stats = ransac<HomographyEstimator, Mat3>(est, ropt, best);
    get_homography_inliers(*best, x1, x2, max_error * max_error, inliers);
    return stats;
}

} // namespace plo
