// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY PARTLY PINNED: sampler, loop control flow, iteration arithmetic, univariate / p3p scalar solvers, Sturm root isolation, F / H scorers, masks, the real-focal check and the scalar camera code against the reference's own code (oracle/_ref, oracle/ref/ref_capi.cc); the transcription of PoseLib's logic for the WHOLE path (solvers, scorers, refiners, estimators, estimate_*) against the reference's own sources run on mini-Eigen (oracle/_ref/libplref2.so, oracle/ref/ref2_capi.cc, tests/test_ref_sources.py); Eigen's own arithmetic (reduction order, decompositions) is UNPINNED.
// Levenberg-Marquardt + normal-equation accumulator + robust losses + the four refiners used by
// LO and by the post-RANSAC polish, restated from PoseLib (paths relative to /root/reference).
#include "plo.h"

namespace plo {
namespace {

// robust/robust_loss.h:41-67,125-136 ; robust_loss.cc:33-56
struct Loss {
    BundleOptions::LossType type;
    double thr, sq_thr, inv_sq_thr;
    explicit Loss(const BundleOptions &opt) : type(opt.loss_type), thr(opt.loss_scale) {
        sq_thr = thr * thr;
        inv_sq_thr = 1.0 / sq_thr;
    }
    double loss(double r2) const {
        switch (type) {
        case BundleOptions::TRUNCATED: return std::min(r2, sq_thr);
        case BundleOptions::HUBER: {
            const double r = std::sqrt(r2);
            return (r <= thr) ? r2 : thr * (2.0 * r - thr);
        }
        case BundleOptions::CAUCHY: return sq_thr * std::log1p(r2 * inv_sq_thr);
        default: return r2;
        }
    }
    double weight(double r2) const {
        switch (type) {
        case BundleOptions::TRUNCATED: return (r2 < sq_thr) ? 1.0 : 0.0;
        case BundleOptions::HUBER: {
            const double r = std::sqrt(r2);
            return (r <= thr) ? 1.0 : thr / r;
        }
        case BundleOptions::CAUCHY: return std::max(std::numeric_limits<double>::min(), 1.0 / (1.0 + r2 * inv_sq_thr));
        default: return 1.0;
        }
    }
};

// robust/optim/jacobian_accumulator.h:46-166.  One residual_count member with the reference's exact
// reset/increment points.
struct NormalAccumulator {
    int np;
    const Loss *loss;
    double residual_acc = 0;
    size_t residual_count = 0;
    double JtJ[8 * 8], Jtr[8];
    NormalAccumulator(int n, const Loss *l) : np(n), loss(l) {
        std::fill(JtJ, JtJ + 64, 0.0);
        std::fill(Jtr, Jtr + 8, 0.0);
    }
    double residual_scale() const { return 1.0 / std::max(1.0, static_cast<double>(residual_count)); }
    void reset_residual() { residual_acc = 0; residual_count = 0; }
    void add_residual1(double res) { residual_acc += loss->loss(res * res); residual_count++; }
    void add_residual2(double r0, double r1) { residual_acc += loss->loss(r0 * r0 + r1 * r1); residual_count++; }
    double get_residual() const { return residual_acc * residual_scale(); }
    void reset_jacobian() {
        residual_count = 0;
        std::fill(JtJ, JtJ + 64, 0.0);
        std::fill(Jtr, Jtr + 8, 0.0);
    }
    // 1-dim residual (:125-141)
    void add_jacobian1(double res, const double *J) {
        const double weight = loss->weight(res * res);
        if (weight == 0) return;
        for (int i = 0; i < np; ++i)
            for (int j = 0; j <= i; ++j) JtJ[i * np + j] += weight * (J[i] * J[j]);
        for (int i = 0; i < np; ++i) Jtr[i] += (weight * res) * J[i];
        residual_count++;
    }
    // 2-dim residual, J is 2 x np row-major (:87-104)
    void add_jacobian2(double r0, double r1, const double *J) {
        const double weight = loss->weight(r0 * r0 + r1 * r1);
        if (weight == 0) return;
        const double *J0 = J, *J1 = J + np;
        for (int i = 0; i < np; ++i)
            for (int j = 0; j <= i; ++j) JtJ[i * np + j] += weight * (J0[i] * J0[j] + J1[i] * J1[j]);
        const double wr0 = weight * r0, wr1 = weight * r1;
        for (int i = 0; i < np; ++i) Jtr[i] += J0[i] * wr0 + J1[i] * wr1;
        residual_count++;
    }
    double grad_norm() const {
        double s = 0;
        for (int i = 0; i < np; ++i) s += Jtr[i] * Jtr[i];
        return residual_scale() * std::sqrt(s);
    }
    // LEVENBERG damping only (types.h:86-89 default) (:145-160)
    void solve(double lambda, double *sol) const {
        const double scale = residual_scale();
        double A[64], rhs[8];
        for (int i = 0; i < np; ++i)
            for (int j = 0; j < np; ++j) A[i * np + j] = scale * JtJ[i * np + j];
        for (int i = 0; i < np; ++i) A[i * np + i] += lambda;
        for (int i = 0; i < np; ++i) rhs[i] = -(scale * Jtr[i]);
        llt_solve_lower(np, A, rhs, sol);
    }
    double predicted_decrease(const double *step, double lambda) const { // :156-160
        const double scale = residual_scale();
        double s = 0;
        for (int i = 0; i < np; ++i) s += step[i] * (lambda * step[i] + scale * Jtr[i]);
        return -s;
    }
};

// robust/optim/lm_impl.h:56-140 (NIELSEN lambda update, no callback)
template <typename Problem, typename Model>
BundleStats lm_impl(Problem &problem, Model *parameters, const BundleOptions &opt) {
    Loss loss(opt);
    BundleStats stats;
    NormalAccumulator acc(problem.num_params, &loss);
    acc.reset_residual();
    stats.cost = problem.compute_residual(acc, *parameters);
    stats.initial_cost = stats.cost;
    stats.grad_norm = -1;
    stats.step_norm = -1;
    stats.invalid_steps = 0;
    stats.lambda = opt.initial_lambda;
    stats.nu = 2.0;
    bool recompute_jac = true;
    double sol[8] = {0};
    for (stats.iterations = 0; stats.iterations < opt.max_iterations; ++stats.iterations) {
        if (recompute_jac) {
            acc.reset_jacobian();
            problem.compute_jacobian(acc, *parameters);
            stats.grad_norm = acc.grad_norm();
            if (stats.grad_norm < opt.gradient_tol) break;
        }
        acc.solve(stats.lambda, sol);
        double sn = 0;
        for (int i = 0; i < problem.num_params; ++i) sn += sol[i] * sol[i];
        stats.step_norm = std::sqrt(sn);
        if (stats.step_norm < opt.step_tol) break;
        Model parameters_new = problem.step(sol, *parameters);
        acc.reset_residual();
        const double cost_new = problem.compute_residual(acc, parameters_new);
        if (cost_new < stats.cost) {
            const double cost_decrease = stats.cost - cost_new;
            *parameters = parameters_new;
            stats.cost = cost_new;
            recompute_jac = true;
            const double predicted = acc.predicted_decrease(sol, stats.lambda);
            if (predicted > 0) {
                const double rho = cost_decrease / predicted;
                const double factor = 1.0 - std::pow(2.0 * rho - 1.0, 3);
                stats.lambda *= std::max(1.0 / 3.0, factor);
            } else {
                stats.lambda *= 1.0 / 3.0;
            }
            stats.nu = 2.0;
            stats.lambda = std::max(opt.min_lambda, stats.lambda);
            if (stats.cost > 0 && cost_decrease / stats.cost < opt.relative_cost_tol) break;
        } else {
            stats.invalid_steps++;
            recompute_jac = false;
            stats.lambda *= stats.nu;
            stats.nu *= 2.0;
            stats.lambda = std::min(opt.max_lambda, stats.lambda);
        }
    }
    return stats;
}

// ---- optim/absolute.h:39-171; NullCameraModel (camera_models.cc:2708-2722) when cam == nullptr,
// PinholeCameraModel (camera_models.cc:668-695) otherwise ------------------------------------------
struct AbsolutePoseRefiner {
    int num_params = 6;
    const std::vector<Vec2> &x;
    const std::vector<Vec3> &X;
    Camera cam; // fixed intrinsics (camera_refine_idx empty); the null camera when none is given
    AbsolutePoseRefiner(const std::vector<Vec2> &x_, const std::vector<Vec3> &X_, const Camera *c)
        : x(x_), X(X_), cam(c ? *c : Camera()) {}
    double compute_residual(NormalAccumulator &acc, const CameraPose &pose) { // absolute.h:49-66
        const Mat3 R = pose.R();
        for (size_t i = 0; i < x.size(); ++i) {
            const Vec3 Z = R * X[i] + pose.t;
            if (Z[2] < 0) continue;
            Vec2 xp;
            cam.project(Z, &xp);
            acc.add_residual2(xp[0] - x[i][0], xp[1] - x[i][1]);
        }
        return acc.get_residual();
    }
    void compute_jacobian(NormalAccumulator &acc, const CameraPose &pose) { // absolute.h:80-130
        const Mat3 R = pose.R();
        for (size_t i = 0; i < x.size(); ++i) {
            const Vec3 Xi = X[i];
            const Vec3 Z = R * Xi + pose.t;
            if (Z[2] < 0) continue;
            Vec2 zp;
            double Jp[2][3];
            cam.project_with_jac(Z, &zp, Jp);
            const double r0 = zp[0] - x[i][0], r1 = zp[1] - x[i][1];
            double dZ[2][3]; // Jproj * R
            for (int a = 0; a < 2; ++a)
                for (int c = 0; c < 3; ++c) dZ[a][c] = Jp[a][0] * R(0, c) + Jp[a][1] * R(1, c) + Jp[a][2] * R(2, c);
            double J[12];
            for (int a = 0; a < 2; ++a) {
                J[a * 6 + 0] = -Xi[2] * dZ[a][1] + Xi[1] * dZ[a][2];
                J[a * 6 + 1] = Xi[2] * dZ[a][0] - Xi[0] * dZ[a][2];
                J[a * 6 + 2] = -Xi[1] * dZ[a][0] + Xi[0] * dZ[a][1];
                J[a * 6 + 3] = dZ[a][0];
                J[a * 6 + 4] = dZ[a][1];
                J[a * 6 + 5] = dZ[a][2];
            }
            acc.add_jacobian2(r0, r1, J);
        }
    }
    CameraPose step(const double *dp, const CameraPose &pose) const {
        CameraPose p;
        p.q = quat_step_post(pose.q, mk3(dp[0], dp[1], dp[2]));
        p.t = pose.t + pose.rotate(mk3(dp[3], dp[4], dp[5]));
        return p;
    }
};

// ---- optim/relative.h:39-166 -----------------------------------------------------------------
// Jacobian of the Sampson residual wrt the 9 entries of E/F (column-major dF[0..8]) (relative.h:129-149)
inline void sampson_resid_and_dF(const Mat3 &E, const Vec2 &p1, const Vec2 &p2, double &r, double dF[9]) {
    const double a0 = p1[0], a1 = p1[1], b0 = p2[0], b1 = p2[1];
    const Vec3 Ex1 = E * mk3(a0, a1, 1.0);
    const double C = b0 * Ex1[0] + b1 * Ex1[1] + 1.0 * Ex1[2];
    double JC[4];
    JC[0] = E(0, 0) * b0 + E(1, 0) * b1 + E(2, 0);
    JC[1] = E(0, 1) * b0 + E(1, 1) * b1 + E(2, 1);
    JC[2] = E(0, 0) * a0 + E(0, 1) * a1 + E(0, 2);
    JC[3] = E(1, 0) * a0 + E(1, 1) * a1 + E(1, 2);
    const double nJC = std::sqrt(JC[0] * JC[0] + JC[1] * JC[1] + JC[2] * JC[2] + JC[3] * JC[3]);
    const double inv_nJC = 1.0 / nJC;
    r = C * inv_nJC;
    dF[0] = a0 * b0; dF[1] = a0 * b1; dF[2] = a0;
    dF[3] = a1 * b0; dF[4] = a1 * b1; dF[5] = a1;
    dF[6] = b0;      dF[7] = b1;      dF[8] = 1.0;
    const double s = C * inv_nJC * inv_nJC;
    dF[0] -= s * (JC[2] * a0 + JC[0] * b0);
    dF[1] -= s * (JC[3] * a0 + JC[0] * b1);
    dF[2] -= s * (JC[0]);
    dF[3] -= s * (JC[2] * a1 + JC[1] * b0);
    dF[4] -= s * (JC[3] * a1 + JC[1] * b1);
    dF[5] -= s * (JC[1]);
    dF[6] -= s * (JC[2]);
    dF[7] -= s * (JC[3]);
    for (int k = 0; k < 9; ++k) dF[k] *= inv_nJC;
}
// residual-only form (relative.h:98-105 / fundamental.h:53-60)
inline double sampson_resid(const Mat3 &E, const Vec2 &p1, const Vec2 &p2) {
    const Vec3 h1 = mk3(p1[0], p1[1], 1.0), h2 = mk3(p2[0], p2[1], 1.0);
    const Vec3 Ex1 = E * h1;
    const double C = dot(h2, Ex1);
    const double n1 = Ex1[0] * Ex1[0] + Ex1[1] * Ex1[1];
    const double t0 = E(0, 0) * h2[0] + E(1, 0) * h2[1] + E(2, 0) * h2[2];
    const double t1 = E(0, 1) * h2[0] + E(1, 1) * h2[1] + E(2, 1) * h2[2];
    const double nJc_sq = n1 + (t0 * t0 + t1 * t1);
    return C / std::sqrt(nJc_sq);
}

// relative.h:62-82
inline void setup_tangent_basis(const Vec3 &t, double tb[3][2]) {
    Vec3 b0;
    const double ax = std::abs(t[0]), ay = std::abs(t[1]), az = std::abs(t[2]);
    if (ax < ay) {
        if (ax < az) b0 = normalized(cross(t, mk3(1, 0, 0)));
        else b0 = normalized(cross(t, mk3(0, 0, 1)));
    } else {
        if (ay < az) b0 = normalized(cross(t, mk3(0, 1, 0)));
        else b0 = normalized(cross(t, mk3(0, 0, 1)));
    }
    const Vec3 b1 = normalized(cross(b0, t));
    for (int r = 0; r < 3; ++r) {
        tb[r][0] = b0[r];
        tb[r][1] = b1[r];
    }
}
// dR (9x3), dt (9x2): derivatives of vec(E) (column-major)  (relative.h:39-60)
inline void deriv_essential_wrt_pose(const Mat3 &E, const Mat3 &R, const double tb[3][2], double dR[9][3],
                                     double dt[9][2]) {
    const Vec3 e0 = col(E, 0), e1 = col(E, 1), e2 = col(E, 2);
    for (int r = 0; r < 3; ++r) {
        dR[r][0] = 0.0;        dR[r][1] = -e2[r];     dR[r][2] = e1[r];
        dR[3 + r][0] = e2[r];  dR[3 + r][1] = 0.0;    dR[3 + r][2] = -e0[r];
        dR[6 + r][0] = -e1[r]; dR[6 + r][1] = e0[r];  dR[6 + r][2] = 0.0;
    }
    const Vec3 tb0 = mk3(tb[0][0], tb[1][0], tb[2][0]), tb1 = mk3(tb[0][1], tb[1][1], tb[2][1]);
    for (int c = 0; c < 3; ++c) {
        const Vec3 v0 = cross(tb0, col(R, c)), v1 = cross(tb1, col(R, c));
        for (int r = 0; r < 3; ++r) {
            dt[3 * c + r][0] = v0[r];
            dt[3 * c + r][1] = v1[r];
        }
    }
}
inline void pose_jac_from_dF(const double dF[9], const double dR[9][3], const double dt[9][2], double J[5]) {
    for (int p = 0; p < 3; ++p) {
        double s = 0;
        for (int m = 0; m < 9; ++m) s += dF[m] * dR[m][p];
        J[p] = s;
    }
    for (int p = 0; p < 2; ++p) {
        double s = 0;
        for (int m = 0; m < 9; ++m) s += dF[m] * dt[m][p];
        J[3 + p] = s;
    }
}

struct RelativePoseRefiner {
    int num_params = 5;
    const std::vector<Vec2> &x1, &x2;
    double tb[3][2]; // tangent basis
    RelativePoseRefiner(const std::vector<Vec2> &a, const std::vector<Vec2> &b) : x1(a), x2(b) {}
    double compute_residual(NormalAccumulator &acc, const CameraPose &pose) {
        Mat3 E;
        essential_from_motion(pose, &E);
        for (size_t k = 0; k < x1.size(); ++k) acc.add_residual1(sampson_resid(E, x1[k], x2[k]));
        return acc.get_residual();
    }
    void compute_jacobian(NormalAccumulator &acc, const CameraPose &pose) {
        const Mat3 R = pose.R();
        Mat3 E;
        essential_from_motion(pose, &E);
        setup_tangent_basis(pose.t, tb);
        double dR[9][3], dt[9][2];
        deriv_essential_wrt_pose(E, R, tb, dR, dt);
        for (size_t k = 0; k < x1.size(); ++k) {
            double r, dF[9], J[5];
            sampson_resid_and_dF(E, x1[k], x2[k], r, dF);
            pose_jac_from_dF(dF, dR, dt, J);
            acc.add_jacobian1(r, J);
        }
    }
    CameraPose step(const double *dp, const CameraPose &pose) const {
        CameraPose p;
        p.q = quat_step_post(pose.q, mk3(dp[0], dp[1], dp[2]));
        for (int r = 0; r < 3; ++r) p.t[r] = pose.t[r] + (tb[r][0] * dp[3] + tb[r][1] * dp[4]);
        return p;
    }
};

// ---- optim/relative.h:167-262 FixCameraRelativePoseRefiner (tangent Sampson error, fixed intrinsics) --------
// Eigen evaluates `M^T * E * d` left to right: the 2x3 product (M^T E) first, then times d.
inline void tangent_terms(const Mat3 &E, const Vec3 &d1, const Vec3 &d2, const Mat32 &M1, const Mat32 &M2, double &C,
                          double JC[4]) {
    const Vec3 Ed1 = E * d1;
    C = d2[0] * Ed1[0] + d2[1] * Ed1[1] + d2[2] * Ed1[2];
    for (int i = 0; i < 2; ++i) {
        double T1[3], T2[3]; // row i of M1^T E^T and of M2^T E
        for (int j = 0; j < 3; ++j) {
            T1[j] = M1.m[0][i] * E(j, 0) + M1.m[1][i] * E(j, 1) + M1.m[2][i] * E(j, 2);
            T2[j] = M2.m[0][i] * E(0, j) + M2.m[1][i] * E(1, j) + M2.m[2][i] * E(2, j);
        }
        JC[i] = T1[0] * d2[0] + T1[1] * d2[1] + T1[2] * d2[2];
        JC[2 + i] = T2[0] * d1[0] + T2[1] * d1[1] + T2[2] * d1[2];
    }
}
struct FixCameraRelativePoseRefiner {
    int num_params = 5;
    const std::vector<Vec3> &d1, &d2;
    const std::vector<Mat32> &M1, &M2;
    double tb[3][2];
    FixCameraRelativePoseRefiner(const std::vector<Vec3> &a, const std::vector<Vec3> &b, const std::vector<Mat32> &m1,
                                 const std::vector<Mat32> &m2)
        : d1(a), d2(b), M1(m1), M2(m2) {}
    double compute_residual(NormalAccumulator &acc, const CameraPose &pose) { // relative.h:181-193
        Mat3 E;
        essential_from_motion(pose, &E);
        for (size_t k = 0; k < d1.size(); ++k) {
            double C, JC[4];
            tangent_terms(E, d1[k], d2[k], M1[k], M2[k], C, JC);
            const double nJc_sq = (JC[2] * JC[2] + JC[3] * JC[3]) + (JC[0] * JC[0] + JC[1] * JC[1]);
            acc.add_residual1(C / std::sqrt(nJc_sq));
        }
        return acc.get_residual();
    }
    void compute_jacobian(NormalAccumulator &acc, const CameraPose &pose) { // relative.h:195-249
        setup_tangent_basis(pose.t, tb);
        const Mat3 R = pose.R();
        Mat3 E;
        essential_from_motion(pose, &E);
        double dR[9][3], dt[9][2];
        deriv_essential_wrt_pose(E, R, tb, dR, dt);
        for (size_t k = 0; k < d1.size(); ++k) {
            const Vec3 &a = d1[k], &b = d2[k];
            const Mat32 &m1 = M1[k], &m2 = M2[k];
            double C, JC[4];
            tangent_terms(E, a, b, m1, m2, C, JC);
            // Vector4d::norm(): packet-of-2 reduction order (SSE2 build) — the one place where this restatement leaves
            // its own left-to-right convention; the reference-order test hook uses the convention
            const double nJ_C = reference_order_enabled()
                                    ? std::sqrt(((JC[0] * JC[0] + JC[1] * JC[1]) + JC[2] * JC[2]) + JC[3] * JC[3])
                                    : std::sqrt((JC[0] * JC[0] + JC[2] * JC[2]) + (JC[1] * JC[1] + JC[3] * JC[3]));
            const double inv_nJ_C = 1.0 / nJ_C;
            const double r = C * inv_nJ_C;
            double dF[9];
            for (int c = 0; c < 3; ++c)
                for (int rr = 0; rr < 3; ++rr) dF[3 * c + rr] = a[c] * b[rr];
            const double s = C * inv_nJ_C * inv_nJ_C;
            for (int c = 0; c < 3; ++c)
                for (int rr = 0; rr < 3; ++rr)
                    dF[3 * c + rr] -= s * (JC[0] * m1.m[c][0] * b[rr] + JC[1] * m1.m[c][1] * b[rr] +
                                           JC[2] * m2.m[rr][0] * a[c] + JC[3] * m2.m[rr][1] * a[c]);
            for (int m = 0; m < 9; ++m) dF[m] *= inv_nJ_C;
            double J[5];
            pose_jac_from_dF(dF, dR, dt, J);
            acc.add_jacobian1(r, J);
        }
    }
    CameraPose step(const double *dp, const CameraPose &pose) const {
        CameraPose p;
        p.q = quat_step_post(pose.q, mk3(dp[0], dp[1], dp[2]));
        for (int r = 0; r < 3; ++r) p.t[r] = pose.t[r] + (tb[r][0] * dp[3] + tb[r][1] * dp[4]);
        return p;
    }
};

// ---- optim/fundamental.h:40-121 + optim_utils.h:57-82 ----------------------------------------
struct FactorizedF {
    Vec4 qU, qV;
    double sigma;
    FactorizedF() {}
    explicit FactorizedF(const Mat3 &F) {
        Mat3 U, V;
        double s[3];
        svd3(F, U, s, V);
        if (det3(U) < 0) U = U * -1.0;
        if (det3(V) < 0) V = V * -1.0;
        qU = rotmat_to_quat(U);
        qV = rotmat_to_quat(V);
        sigma = s[1] / s[0];
    }
    Mat3 F() const {
        const Mat3 U = quat_to_rotmat(qU), V = quat_to_rotmat(qV);
        Mat3 out;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) out(r, c) = U(r, 0) * V(c, 0) + sigma * U(r, 1) * V(c, 1);
        return out;
    }
};
struct FundamentalRefiner {
    int num_params = 7;
    const std::vector<Vec2> &x1, &x2;
    FundamentalRefiner(const std::vector<Vec2> &a, const std::vector<Vec2> &b) : x1(a), x2(b) {}
    double compute_residual(NormalAccumulator &acc, const FactorizedF &FF) {
        const Mat3 F = FF.F();
        for (size_t k = 0; k < x1.size(); ++k) acc.add_residual1(sampson_resid(F, x1[k], x2[k]));
        return acc.get_residual();
    }
    void compute_jacobian(NormalAccumulator &acc, const FactorizedF &FF) {
        const Mat3 F = FF.F();
        const Mat3 U = quat_to_rotmat(FF.qU), V = quat_to_rotmat(FF.qV);
        // dF/dparams (9x7, vec column-major): U' = exp([wU]x) U -> dF = [e_k]x F ;
        // V' = exp([wV]x) V -> dF = -F [e_k]x ; sigma -> u1 v1^T     (fundamental.h:68-77)
        double D[9][7];
        for (int cidx = 0; cidx < 3; ++cidx)
            for (int r = 0; r < 3; ++r) {
                const int m = 3 * cidx + r;
                // ([e_k]x F)(:,c): k=0: (0, -F2, F1) ; k=1: (F2, 0, -F0) ; k=2: (-F1, F0, 0)
                const double f0 = F(0, cidx), f1 = F(1, cidx), f2 = F(2, cidx);
                const double gx[3][3] = {{0.0, -f2, f1}, {f2, 0.0, -f0}, {-f1, f0, 0.0}};
                for (int k = 0; k < 3; ++k) D[m][k] = gx[k][r];
                // (-F [e_k]x)(r,:): k=0: (0, -Fr2, Fr1) ; k=1: (Fr2, 0, -Fr0) ; k=2: (-Fr1, Fr0, 0)
                const double fr0 = F(r, 0), fr1 = F(r, 1), fr2 = F(r, 2);
                const double hx[3][3] = {{0.0, -fr2, fr1}, {fr2, 0.0, -fr0}, {-fr1, fr0, 0.0}};
                for (int k = 0; k < 3; ++k) D[m][3 + k] = hx[k][cidx];
                D[m][6] = U(r, 1) * V(cidx, 1);
            }
        for (size_t k = 0; k < x1.size(); ++k) {
            double r, dF[9], J[7];
            sampson_resid_and_dF(F, x1[k], x2[k], r, dF);
            for (int p = 0; p < 7; ++p) {
                double s = 0;
                for (int m = 0; m < 9; ++m) s += dF[m] * D[m][p];
                J[p] = s;
            }
            acc.add_jacobian1(r, J);
        }
    }
    FactorizedF step(const double *dp, const FactorizedF &F) const {
        FactorizedF n;
        n.qU = quat_step_pre(F.qU, mk3(dp[0], dp[1], dp[2]));
        n.qV = quat_step_pre(F.qV, mk3(dp[3], dp[4], dp[5]));
        n.sigma = F.sigma + dp[6];
        return n;
    }
};

// ---- optim/homography.h:45-178 ---------------------------------------------------------------
inline Mat3 adjugate(const Mat3 &H) { // homography.h:160-175
    Mat3 a;
    a(0, 0) = H(1, 1) * H(2, 2) - H(1, 2) * H(2, 1);
    a(0, 1) = H(0, 2) * H(2, 1) - H(0, 1) * H(2, 2);
    a(0, 2) = H(0, 1) * H(1, 2) - H(0, 2) * H(1, 1);
    a(1, 0) = H(1, 2) * H(2, 0) - H(1, 0) * H(2, 2);
    a(1, 1) = H(0, 0) * H(2, 2) - H(0, 2) * H(2, 0);
    a(1, 2) = H(0, 2) * H(1, 0) - H(0, 0) * H(1, 2);
    a(2, 0) = H(1, 0) * H(2, 1) - H(1, 1) * H(2, 0);
    a(2, 1) = H(0, 1) * H(2, 0) - H(0, 0) * H(2, 1);
    a(2, 2) = H(0, 0) * H(1, 1) - H(0, 1) * H(1, 0);
    return a;
}
struct HomographyRefiner {
    int num_params = 8;
    const std::vector<Vec2> &x1, &x2;
    HomographyRefiner(const std::vector<Vec2> &a, const std::vector<Vec2> &b) : x1(a), x2(b) {}
    double compute_residual(NormalAccumulator &acc, const Mat3 &H) {
        const Mat3 G = adjugate(H);
        for (size_t k = 0; k < x1.size(); ++k) {
            const double a0 = x1[k][0], a1 = x1[k][1], b0 = x2[k][0], b1 = x2[k][1];
            const double Hx0 = H(0, 0) * a0 + H(0, 1) * a1 + H(0, 2);
            const double Hx1 = H(1, 0) * a0 + H(1, 1) * a1 + H(1, 2);
            const double iw = 1.0 / (H(2, 0) * a0 + H(2, 1) * a1 + H(2, 2));
            acc.add_residual2(Hx0 * iw - b0, Hx1 * iw - b1);
            const double Gx0 = G(0, 0) * b0 + G(0, 1) * b1 + G(0, 2);
            const double Gx1 = G(1, 0) * b0 + G(1, 1) * b1 + G(1, 2);
            const double iv = 1.0 / (G(2, 0) * b0 + G(2, 1) * b1 + G(2, 2));
            acc.add_residual2(Gx0 * iv - a0, Gx1 * iv - a1);
        }
        return acc.get_residual();
    }
    void compute_jacobian(NormalAccumulator &acc, const Mat3 &H) {
        const Mat3 G = adjugate(H);
        for (size_t k = 0; k < x1.size(); ++k) {
            const double a0 = x1[k][0], a1 = x1[k][1], b0 = x2[k][0], b1 = x2[k][1];
            // forward block (homography.h:107-123); params = first 8 column-major entries of H
            const double Hx0 = H(0, 0) * a0 + H(0, 1) * a1 + H(0, 2);
            const double Hx1 = H(1, 0) * a0 + H(1, 1) * a1 + H(1, 2);
            const double iw = 1.0 / (H(2, 0) * a0 + H(2, 1) * a1 + H(2, 2));
            const double z0 = Hx0 * iw, z1 = Hx1 * iw;
            double J[16] = {a0, 0.0, -a0 * z0, a1, 0.0, -a1 * z0, 1.0, 0.0,
                            0.0, a0, -a0 * z1, 0.0, a1, -a1 * z1, 0.0, 1.0};
            for (int m = 0; m < 16; ++m) J[m] = J[m] * iw;
            acc.add_jacobian2(z0 - b0, z1 - b1, J);
            // backward block (homography.h:125-152): y = pi(adj(H) x2); chain rule through adj(H)
            const double Gx0 = G(0, 0) * b0 + G(0, 1) * b1 + G(0, 2);
            const double Gx1 = G(1, 0) * b0 + G(1, 1) * b1 + G(1, 2);
            const double iv = 1.0 / (G(2, 0) * b0 + G(2, 1) * b1 + G(2, 2));
            const double y0 = Gx0 * iv, y1 = Gx1 * iv;
            const double y0b1 = y0 * b1, y0b0 = y0 * b0, y1b1 = y1 * b1, y1b0 = y1 * b0;
            const double H0_0 = H(0, 0), H0_1 = H(0, 1), H0_2 = H(0, 2);
            const double H1_0 = H(1, 0), H1_1 = H(1, 1), H1_2 = H(1, 2);
            const double H2_0 = H(2, 0), H2_1 = H(2, 1), H2_2 = H(2, 2);
            // d(G x2 - y * (G x2)_2)/dH_k for k over (H00,H10,H20,H01,H11,H21,H02,H12)
            double Jb[16];
            // row 0 (y0)
            Jb[0] = H2_1 * y0b1 - H1_1 * y0;                         // dH00
            Jb[1] = H0_1 * y0 - H2_1 * y0b0;                         // dH10
            Jb[2] = H1_1 * y0b0 - H0_1 * y0b1;                       // dH20
            Jb[3] = H1_2 - H2_2 * b1 + H1_0 * y0 - H2_0 * y0b1;      // dH01
            Jb[4] = H2_2 * b0 - H0_2 - H0_0 * y0 + H2_0 * y0b0;      // dH11
            Jb[5] = H0_2 * b1 - H1_2 * b0 + H0_0 * y0b1 - H1_0 * y0b0; // dH21
            Jb[6] = H2_1 * b1 - H1_1;                                // dH02
            Jb[7] = H0_1 - H2_1 * b0;                                // dH12
            // row 1 (y1)
            Jb[8] = H2_2 * b1 - H1_2 - H1_1 * y1 + H2_1 * y1b1;      // dH00
            Jb[9] = H0_2 - H2_2 * b0 + H0_1 * y1 - H2_1 * y1b0;      // dH10
            Jb[10] = H1_2 * b0 - H0_2 * b1 - H0_1 * y1b1 + H1_1 * y1b0; // dH20
            Jb[11] = H1_0 * y1 - H2_0 * y1b1;                        // dH01
            Jb[12] = H2_0 * y1b0 - H0_0 * y1;                        // dH11
            Jb[13] = H0_0 * y1b1 - H1_0 * y1b0;                      // dH21
            Jb[14] = H1_0 - H2_0 * b1;                               // dH02
            Jb[15] = H2_0 * b0 - H0_0;                               // dH12
            for (int m = 0; m < 16; ++m) Jb[m] = Jb[m] * iv;
            acc.add_jacobian2(y0 - a0, y1 - a1, Jb);
        }
    }
    Mat3 step(const double *dp, const Mat3 &H) const { // first 8 column-major entries
        Mat3 n = H;
        for (int m = 0; m < 8; ++m) n(m % 3, m / 3) += dp[m];
        return n;
    }
};

} // namespace

// bundle.cc:84-112 (calibrated interface: NullCameraModel; `weights` ignored by the reference, :89)
BundleStats bundle_adjust(const std::vector<Vec2> &x, const std::vector<Vec3> &X, CameraPose *pose,
                          const BundleOptions &opt) {
    AbsolutePoseRefiner refiner(x, X, nullptr);
    return lm_impl(refiner, pose, opt);
}
// bundle.cc:95-112 with a fixed camera (no intrinsics refinement)
BundleStats bundle_adjust_camera(const std::vector<Vec2> &x, const std::vector<Vec3> &X, const Camera &cam,
                                 CameraPose *pose, const BundleOptions &opt) {
    AbsolutePoseRefiner refiner(x, X, &cam);
    return lm_impl(refiner, pose, opt);
}
// bundle.cc:206-222
BundleStats refine_relpose(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, CameraPose *pose,
                           const BundleOptions &opt) {
    RelativePoseRefiner refiner(x1, x2);
    return lm_impl(refiner, pose, opt);
}
// bundle.cc:227-235,268-277 (pre-computed bearings and unprojection Jacobians, uniform weights)
BundleStats refine_relpose(const std::vector<Vec3> &d1, const std::vector<Vec3> &d2, const std::vector<Mat32> &M1,
                           const std::vector<Mat32> &M2, CameraPose *pose, const BundleOptions &opt) {
    FixCameraRelativePoseRefiner refiner(d1, d2, M1, M2);
    return lm_impl(refiner, pose, opt);
}
// bundle.cc:313-333
BundleStats refine_fundamental(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, Mat3 *F,
                               const BundleOptions &opt) {
    FactorizedF FF(*F);
    FundamentalRefiner refiner(x1, x2);
    BundleStats stats = lm_impl(refiner, &FF, opt);
    *F = FF.F();
    return stats;
}
// bundle.cc:394-411
BundleStats refine_homography(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, Mat3 *H,
                              const BundleOptions &opt) {
    HomographyRefiner refiner(x1, x2);
    return lm_impl(refiner, H, opt);
}

} // namespace plo
