// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY PARTLY PINNED: sampler, loop control flow, iteration arithmetic, univariate / p3p scalar solvers, Sturm root isolation, F / H scorers, masks, the real-focal check and the scalar camera code against the reference's own code (oracle/_ref, oracle/ref/ref_capi.cc); the transcription of PoseLib's logic for the WHOLE path (solvers, scorers, refiners, estimators, estimate_*) against the reference's own sources run on mini-Eigen (oracle/_ref/libplref2.so, oracle/ref/ref2_capi.cc, tests/test_ref_sources.py); Eigen's own arithmetic (reduction order, decompositions) is UNPINNED.
// Minimal solvers + their scalar helpers, restated from PoseLib (paths relative to /root/reference).
#include "plo.h"
#include <array>

namespace plo {

// ============================ misc/univariate.cc ==============================================
// univariate.cc:48-61
int solve_quadratic_real(double a, double b, double c, double roots[2]) {
    const double disc = b * b - 4 * a * c;
    if (disc < 0) return 0;
    const double sq = std::sqrt(disc);
    roots[0] = (b > 0) ? (2 * c) / (-b - sq) : (2 * c) / (-b + sq);
    roots[1] = c / (a * roots[0]);
    return 2;
}
// univariate.cc:74-92
bool solve_cubic_single_real(double c2, double c1, double c0, double &root) {
    double a = c1 - c2 * c2 / 3.0;
    double b = (2.0 * c2 * c2 * c2 - 9.0 * c2 * c1) / 27.0 + c0;
    double c = b * b / 4.0 + a * a * a / 27.0;
    if (c != 0) {
        if (c > 0) {
            c = std::sqrt(c);
            b *= -0.5;
            root = std::cbrt(b + c) + std::cbrt(b - c) - c2 / 3.0;
            return true;
        } else {
            c = 3.0 * b / (2.0 * a) * std::sqrt(-3.0 / a);
            root = 2.0 * std::sqrt(-a / 3.0) * std::cos(std::acos(c) / 3.0) - c2 / 3.0;
        }
    } else {
        root = -c2 / 3.0 + (a != 0 ? (3.0 * b / a) : 0);
    }
    return false;
}
// univariate.cc:94-126
int solve_cubic_real(double c2, double c1, double c0, double roots[3]) {
    double a = c1 - c2 * c2 / 3.0;
    double b = (2.0 * c2 * c2 * c2 - 9.0 * c2 * c1) / 27.0 + c0;
    double c = b * b / 4.0 + a * a * a / 27.0;
    int n_roots;
    if (a == 0.0 && b == 0.0) {
        roots[0] = roots[1] = roots[2] = -c2 / 3.0;
        n_roots = 3;
    } else if (c > 0) {
        c = std::sqrt(c);
        b *= -0.5;
        roots[0] = std::cbrt(b + c) + std::cbrt(b - c) - c2 / 3.0;
        n_roots = 1;
    } else {
        c = 3.0 * b / (2.0 * a) * std::sqrt(-3.0 / a);
        const double d = 2.0 * std::sqrt(-a / 3.0);
        const double third = std::acos(c) / 3.0;
        roots[0] = d * std::cos(third) - c2 / 3.0;
        roots[1] = d * std::cos(third - 2.09439510239319526263557236234192) - c2 / 3.0;
        roots[2] = d * std::cos(third - 4.18879020478639052527114472468384) - c2 / 3.0;
        n_roots = 3;
    }
    for (int i = 0; i < n_roots; ++i) { // one Newton step per root (:117-124)
        const double x = roots[i], x2 = x * x, x3 = x * x2;
        const double dx = -(x3 + c2 * x2 + c1 * x + c0) / (3 * x2 + 2 * c2 * x + c1);
        roots[i] += dx;
    }
    return n_roots;
}

// ============================ misc/sturm.h (N = 10 instantiation) =============================
namespace {
constexpr int SN = 10;
constexpr int STURM_DEPTH_LIMIT = 300; // CMakeLists.txt:48-52 MAX_STURM_RECURSION_DEPTH_LIMIT

// sturm.h:47-84.  fvec = [monic poly (N+1) | monic-normalised derivative (N)], svec = 3N quotients.
void build_sturm_seq(const double *fvec, double *svec) {
    double f[3 * SN];
    double *f1 = f, *f2 = f1 + SN + 1, *f3 = f2 + SN;
    std::copy(fvec, fvec + (2 * SN + 1), f);
    for (int i = 0; i < SN - 1; ++i) {
        const double q1 = f1[SN - i] * f2[SN - 1 - i];
        const double q0 = f1[SN - 1 - i] * f2[SN - 1 - i] - f1[SN - i] * f2[SN - 2 - i];
        f3[0] = f1[0] - q0 * f2[0];
        for (int j = 1; j < SN - 1 - i; ++j) f3[j] = f1[j] - q1 * f2[j - 1] - q0 * f2[j];
        const double c = -std::abs(f3[SN - 2 - i]);
        const double ci = 1.0 / c;
        for (int j = 0; j < SN - 1 - i; ++j) f3[j] = f3[j] * ci;
        double *tmp = f1;
        f1 = f2;
        f2 = f3;
        f3 = tmp;
        svec[3 * i] = q0;
        svec[3 * i + 1] = q1;
        svec[3 * i + 2] = c;
    }
    svec[3 * SN - 3] = f1[0];
    svec[3 * SN - 2] = f1[1];
    svec[3 * SN - 1] = f2[0];
}
// sturm.h:88-94 (monic, degree deg)
inline double polyval_monic(const double *f, int deg, double x) {
    double fx = x + f[deg - 1];
    for (int i = deg - 2; i >= 0; --i) fx = x * fx + f[i];
    return fx;
}
// sturm.h:102-117
int signchanges(const double *svec, double x) {
    double f[SN + 1];
    f[SN] = svec[3 * SN - 1];
    f[SN - 1] = svec[3 * SN - 3] + x * svec[3 * SN - 2];
    for (int i = SN - 2; i >= 0; --i) f[i] = (svec[3 * i] + x * svec[3 * i + 1]) * f[i + 1] + svec[3 * i + 2] * f[i + 2];
    int count = 0;
    for (int i = 0; i < SN; ++i)
        if ((f[i] < 0) != (f[i + 1] < 0)) ++count;
    return count;
}
// sturm.h:153-208
void ridders_method_newton(const double *fvec, double a, double b, double *roots, int &n_roots, double tol) {
    double fa = polyval_monic(fvec, SN, a);
    double fb = polyval_monic(fvec, SN, b);
    if (!((fa < 0) ^ (fb < 0))) return;
    const double tol_newton = 1e-3;
    for (int iter = 0; iter < 30; ++iter) {
        if (std::abs(a - b) < tol_newton) break;
        const double c = (a + b) * 0.5;
        const double fc = polyval_monic(fvec, SN, c);
        const double s = std::sqrt(fc * fc - fa * fb);
        if (!s) break;
        const double d = (fa < fb) ? c + (a - c) * fc / s : c + (c - a) * fc / s;
        const double fd = polyval_monic(fvec, SN, d);
        if (fd >= 0 ? (fc < 0) : (fc > 0)) {
            a = c; fa = fc; b = d; fb = fd;
        } else if (fd >= 0 ? (fa < 0) : (fa > 0)) {
            b = d; fb = fd;
        } else {
            a = d; fa = fd;
        }
    }
    double x = (a + b) * 0.5;
    const double *fpvec = fvec + SN + 1;
    for (int iter = 0; iter < 10; ++iter) {
        const double fx = polyval_monic(fvec, SN, x);
        if (std::abs(fx) < tol) break;
        const double fpx = double(SN) * polyval_monic(fpvec, SN - 1, x);
        const double dx = fx / fpx;
        x = x - dx;
        if (std::abs(dx) < tol) break;
    }
    roots[n_roots++] = x;
}
// sturm.h:210-231
void isolate_roots(const double *fvec, const double *svec, double a, double b, int sa, int sb, double *roots,
                   int &n_roots, double tol, int depth) {
    if (depth > STURM_DEPTH_LIMIT) return;
    if (b - a < tol) {
        roots[n_roots++] = b;
        return;
    }
    const int n_rts = sa - sb;
    if (n_rts > 1) {
        const double c = (a + b) * 0.5;
        const int sc = signchanges(svec, c);
        isolate_roots(fvec, svec, a, c, sa, sc, roots, n_roots, tol, depth + 1);
        isolate_roots(fvec, svec, c, b, sc, sb, roots, n_roots, tol, depth + 1);
    } else if (n_rts == 1) {
        ridders_method_newton(fvec, a, b, roots, n_roots, tol);
    }
}
} // namespace

// sturm.h:233-274
int bisect_sturm10(const double *coeffs, double *roots, double tol) {
    if (coeffs[SN] == 0.0) return 0;
    double fvec[2 * SN + 1], svec[3 * SN];
    std::copy(coeffs, coeffs + SN + 1, fvec);
    const double c_inv = 1.0 / fvec[SN];
    for (int i = 0; i < SN; ++i) fvec[i] *= c_inv;
    fvec[SN] = 1.0;
    for (int i = 0; i < SN - 1; ++i) fvec[SN + 1 + i] = fvec[i + 1] * ((i + 1) / double(SN));
    fvec[2 * SN] = 1.0;
    build_sturm_seq(fvec, svec);
    double mx = 0; // get_bounds :144-150
    for (int i = 0; i < SN; ++i) mx = std::max(mx, std::abs(fvec[i]));
    const double r0 = 1.0 + mx;
    const double a = -r0, b = r0;
    const int sa = signchanges(svec, a), sb = signchanges(svec, b);
    int n_roots = sa - sb;
    if (n_roots == 0) return 0;
    n_roots = 0;
    isolate_roots(fvec, svec, a, b, sa, sb, roots, n_roots, tol, 0);
    return n_roots;
}

// ============================ misc/essential.cc ===============================================
// essential.cc:35-38   E = [t]x R
void essential_from_motion(const CameraPose &pose, Mat3 *E) {
    Mat3 Tx;
    Tx(0, 0) = 0.0;        Tx(0, 1) = -pose.t[2]; Tx(0, 2) = pose.t[1];
    Tx(1, 0) = pose.t[2];  Tx(1, 1) = 0.0;        Tx(1, 2) = -pose.t[0];
    Tx(2, 0) = -pose.t[1]; Tx(2, 1) = pose.t[0];  Tx(2, 2) = 0.0;
    *E = Tx * pose.R();
}
// essential.cc:40-57
bool check_cheirality(const CameraPose &pose, const Vec3 &x1, const Vec3 &x2, double min_depth) {
    const Vec3 Rx1 = pose.rotate(x1);
    const double a = -dot(Rx1, x2);
    const double b1 = -dot(Rx1, pose.t);
    const double b2 = dot(x2, pose.t);
    const double lambda1 = b1 - a * b2;
    const double lambda2 = -a * b1 + b2;
    min_depth = min_depth * (1 - a * a);
    return lambda1 > min_depth && lambda2 > min_depth;
}
static bool check_cheirality_all(const CameraPose &pose, const std::vector<Vec3> &x1, const std::vector<Vec3> &x2) {
    for (size_t i = 0; i < x1.size(); ++i) // essential.cc:81-89
        if (!check_cheirality(pose, x1[i], x2[i], 0.0)) return false;
    return true;
}
// essential.cc:103-169
void motion_from_essential(const Mat3 &E, const std::vector<Vec3> &x1, const std::vector<Vec3> &x2,
                           std::vector<CameraPose> *poses) {
    const Vec3 e0 = col(E, 0), e1 = col(E, 1), e2 = col(E, 2);
    const Vec3 u12 = cross(e0, e1), u13 = cross(e0, e2), u23 = cross(e1, e2);
    const double n12 = sqnorm(u12), n13 = sqnorm(u13), n23 = sqnorm(u23);
    Mat3 UW, Vt;
    if (n12 > n13) {
        if (n12 > n23) {
            set_col(UW, 1, normalized(e0));
            set_col(UW, 2, u12 / std::sqrt(n12));
        } else {
            set_col(UW, 1, normalized(e1));
            set_col(UW, 2, u23 / std::sqrt(n23));
        }
    } else {
        if (n13 > n23) {
            set_col(UW, 1, normalized(e0));
            set_col(UW, 2, u13 / std::sqrt(n13));
        } else {
            set_col(UW, 1, normalized(e1));
            set_col(UW, 2, u23 / std::sqrt(n23));
        }
    }
    set_col(UW, 0, -cross(col(UW, 2), col(UW, 1)));
    const Mat3 Et = transpose(E);
    Vec3 v0 = Et * col(UW, 1);      // (UW.col(1)^T E)^T
    Vec3 v1 = Et * (-col(UW, 0));   // (-UW.col(0)^T E)^T
    v0 = normalized(v0);
    v1 = v1 - dot(v0, v1) * v0;
    v1 = normalized(v1);
    set_row(Vt, 0, v0);
    set_row(Vt, 1, v1);
    set_row(Vt, 2, cross(v0, v1));

    CameraPose pose;
    pose.q = rotmat_to_quat(UW * Vt);
    pose.t = col(UW, 2);
    if (check_cheirality_all(pose, x1, x2)) poses->push_back(pose);
    pose.t = -pose.t;
    if (check_cheirality_all(pose, x1, x2)) poses->push_back(pose);
    for (int r = 0; r < 3; ++r) { // U * W^T: negate first two columns
        UW(r, 0) = -UW(r, 0);
        UW(r, 1) = -UW(r, 1);
    }
    pose.q = rotmat_to_quat(UW * Vt);
    if (check_cheirality_all(pose, x1, x2)) poses->push_back(pose);
    pose.t = -pose.t;
    if (check_cheirality_all(pose, x1, x2)) poses->push_back(pose);
}

// ============================ solvers/p3p.cc + p3p_common.h ===================================
namespace {
// p3p_common.h:7-29
bool root2real(double b, double c, double &r1, double &r2) {
    const double THRESHOLD = -1.0e-12;
    const double v = b * b - 4.0 * c;
    if (v < THRESHOLD) {
        r1 = r2 = -0.5 * b;
        return v >= 0;
    }
    if (v > THRESHOLD && v < 0.0) {
        r1 = -0.5 * b;
        r2 = -2;
        return true;
    }
    const double y = std::sqrt(v);
    if (b < 0) {
        r1 = 0.5 * (-b + y);
        r2 = 0.5 * (-b - y);
    } else {
        r1 = 2.0 * c / (-b + y);
        r2 = 2.0 * c / (-b - y);
    }
    return true;
}
// p3p_common.h:31-70 — split the degenerate conic C into its two lines p, q
void compute_pq(Mat3 C, Vec3 pq[2]) {
    Mat3 A; // adjugate-like matrix of the symmetric C (sign convention of the reference)
    A(0, 0) = C(1, 2) * C(2, 1) - C(1, 1) * C(2, 2);
    A(1, 1) = C(0, 2) * C(2, 0) - C(0, 0) * C(2, 2);
    A(2, 2) = C(0, 1) * C(1, 0) - C(0, 0) * C(1, 1);
    A(0, 1) = C(0, 1) * C(2, 2) - C(0, 2) * C(2, 1);
    A(0, 2) = C(0, 2) * C(1, 1) - C(0, 1) * C(1, 2);
    A(1, 0) = A(0, 1);
    A(1, 2) = C(0, 0) * C(1, 2) - C(0, 2) * C(1, 0);
    A(2, 0) = A(0, 2);
    A(2, 1) = A(1, 2);
    Vec3 v;
    if (A(0, 0) > A(1, 1)) {
        if (A(0, 0) > A(2, 2)) v = col(A, 0) / std::sqrt(A(0, 0));
        else v = col(A, 2) / std::sqrt(A(2, 2));
    } else if (A(1, 1) > A(2, 2)) {
        v = col(A, 1) / std::sqrt(A(1, 1));
    } else {
        v = col(A, 2) / std::sqrt(A(2, 2));
    }
    C(0, 1) -= v[2]; C(0, 2) += v[1]; C(1, 2) -= v[0];
    C(1, 0) += v[2]; C(2, 0) -= v[1]; C(2, 1) += v[0];
    pq[0] = col(C, 0);
    pq[1] = row(C, 0);
}
// p3p_common.h:72-94 — Newton on the three distance equations
void refine_lambda(double &l1, double &l2, double &l3, double a12, double a13, double a23, double b12, double b13,
                   double b23) {
    for (int iter = 0; iter < 5; ++iter) {
        const double r1 = (l1 * l1 - 2.0 * l1 * l2 * b12 + l2 * l2 - a12);
        const double r2 = (l1 * l1 - 2.0 * l1 * l3 * b13 + l3 * l3 - a13);
        const double r3 = (l2 * l2 - 2.0 * l2 * l3 * b23 + l3 * l3 - a23);
        if (std::abs(r1) + std::abs(r2) + std::abs(r3) < 1e-10) return;
        const double x11 = l1 - l2 * b12, x12 = l2 - l1 * b12;
        const double x21 = l1 - l3 * b13, x23 = l3 - l1 * b13;
        const double x32 = l2 - l3 * b23, x33 = l3 - l2 * b23;
        const double detJ = 0.5 / (x11 * x23 * x32 + x12 * x21 * x33);
        l1 += (-x23 * x32 * r1 - x12 * x33 * r2 + x12 * x23 * r3) * detJ;
        l2 += (-x21 * x33 * r1 + x11 * x33 * r2 - x11 * x23 * r3) * detJ;
        l3 += (x21 * x32 * r1 - x11 * x32 * r2 - x12 * x21 * r3) * detJ;
    }
}
} // namespace

// p3p.cc:39-202
int p3p(const std::vector<Vec3> &x_in, const std::vector<Vec3> &X_in, std::vector<CameraPose> *output) {
    if (output == nullptr) return 0;
    output->clear();
    Vec3 X01 = X_in[0] - X_in[1], X02 = X_in[0] - X_in[2], X12 = X_in[1] - X_in[2];
    double a01 = sqnorm(X01), a02 = sqnorm(X02), a12 = sqnorm(X12);
    std::array<Vec3, 3> X = {X_in[0], X_in[1], X_in[2]};
    std::array<Vec3, 3> x = {x_in[0], x_in[1], x_in[2]};
    // reorder so that |X1-X2| is the largest distance (:58-73)
    if (a01 > a02) {
        if (a01 > a12) {
            std::swap(x[0], x[2]);
            std::swap(X[0], X[2]);
            std::swap(a01, a12);
            X01 = -X12;
            X02 = -X02;
        }
    } else if (a02 > a12) {
        std::swap(x[0], x[1]);
        std::swap(X[0], X[1]);
        std::swap(a02, a12);
        X01 = -X01;
        X02 = X12;
    }
    const double a12d = 1.0 / a12;
    const double a = a01 * a12d, b = a02 * a12d;
    const double m01 = dot(x[0], x[1]), m02 = dot(x[0], x[2]), m12 = dot(x[1], x[2]);
    // :84-98
    const double m12sq = -m12 * m12 + 1.0;
    const double m02sq = -1.0 + m02 * m02;
    const double m01sq = -1.0 + m01 * m01;
    const double ab = a * b, bsq = b * b, asq = a * a;
    const double m013 = -2.0 + 2.0 * m01 * m02 * m12;
    const double bsqm12sq = bsq * m12sq, asqm12sq = asq * m12sq, abm12sq = 2.0 * ab * m12sq;
    const double k3_inv = 1.0 / (bsqm12sq + b * m02sq);
    const double k2 = k3_inv * ((-1.0 + a) * m02sq + abm12sq + bsqm12sq + b * m013);
    const double k1 = k3_inv * (asqm12sq + abm12sq + a * m013 + (-1.0 + b) * m01sq);
    const double k0 = k3_inv * (asqm12sq + a * m01sq);
    double s;
    const bool G = solve_cubic_single_real(k2, k1, k0, s);
    Mat3 C; // :103-112
    C(0, 0) = -a + s * (1 - b);
    C(0, 1) = -m02 * s;
    C(0, 2) = a * m12 + b * m12 * s;
    C(1, 0) = C(0, 1);
    C(1, 1) = s + 1;
    C(1, 2) = -m01;
    C(2, 0) = C(0, 2);
    C(2, 1) = C(1, 2);
    C(2, 2) = -a - b * s + 1;
    Vec3 pq[2];
    compute_pq(C, pq);
    output->clear();
    Mat3 XX;
    set_col(XX, 0, X01);
    set_col(XX, 1, X02);
    set_col(XX, 2, cross(X01, X02));
    XX = inverse3(XX);
    int n_sols = 0;
    auto emit = [&](double d0, double d1, double d2) {
        refine_lambda(d0, d1, d2, a01, a02, a12, m01, m02, m12);
        const Vec3 v1 = d0 * x[0] - d1 * x[1];
        const Vec3 v2 = d0 * x[0] - d2 * x[2];
        Mat3 YY;
        set_col(YY, 0, v1);
        set_col(YY, 1, v2);
        set_col(YY, 2, cross(v1, v2));
        const Mat3 R = YY * XX;
        output->push_back(pose_from_Rt(R, d0 * x[0] - R * X[0]));
        ++n_sols;
    };
    for (int i = 0; i < 2; ++i) {
        const double p0 = pq[i][0], p1 = pq[i][1], p2 = pq[i][2];
        const bool switch_12 = std::abs(p0) <= std::abs(p1);
        if (switch_12) { // eliminate d0 (:136-168)
            const double w0 = -p0 / p1, w1 = -p2 / p1;
            const double ca = 1.0 / (w1 * w1 - b);
            const double cb = 2.0 * (b * m12 - m02 * w1 + w0 * w1) * ca;
            const double cc = (w0 * w0 - 2 * m02 * w0 - b + 1.0) * ca;
            double taus[2];
            if (!root2real(cb, cc, taus[0], taus[1])) continue;
            for (double tau : taus) {
                if (tau <= 0) continue;
                const double d2 = std::sqrt(a12 / (tau * (tau - 2.0 * m12) + 1.0));
                const double d1 = tau * d2;
                const double d0 = (w0 * d2 + w1 * d1);
                if (d0 < 0) continue;
                emit(d0, d1, d2);
            }
        } else { // (:169-197)
            const double w0 = -p1 / p0, w1 = -p2 / p0;
            const double ca = 1.0 / (-a * w1 * w1 + 2 * a * m12 * w1 - a + 1);
            const double cb = 2 * (a * m12 * w0 - m01 - a * w0 * w1) * ca;
            const double cc = (1 - a * w0 * w0) * ca;
            double taus[2];
            if (!root2real(cb, cc, taus[0], taus[1])) continue;
            for (double tau : taus) {
                if (tau <= 0) continue;
                const double d0 = std::sqrt(a01 / (tau * (tau - 2.0 * m01) + 1.0));
                const double d1 = tau * d0;
                const double d2 = w0 * d0 + w1 * d1;
                if (d2 < 0) continue;
                emit(d0, d1, d2);
            }
        }
        if (n_sols > 0 && G) break;
    }
    return (int)output->size();
}

// ============================ solvers/p3p_lambdatwist.cc =====================================
// Persson & Nordberg's Lambda Twist P3P as PoseLib implements it (an alternative to p3p above).
// Eigen's reductions (dot, squaredNorm, array sum) are taken left to right / column-major, as everywhere in this file.
namespace {
// p3p_lambdatwist.cc:36-65: the two non-zero eigenvalues (|sig1| >= |sig2|) and their eigenvectors (columns e1, e2)
// of a symmetric 3x3 matrix that has a zero eigenvalue.
void eig3x3_known0(const Mat3 &M, Vec3 &e1, Vec3 &e2, double &sig1, double &sig2) {
    const double p1 = -M(0, 0) - M(1, 1) - M(2, 2);
    const double p0 =
        -M(0, 1) * M(0, 1) - M(0, 2) * M(0, 2) - M(1, 2) * M(1, 2) + M(0, 0) * (M(1, 1) + M(2, 2)) + M(1, 1) * M(2, 2);
    const double disc = std::sqrt(p1 * p1 / 4.0 - p0);
    const double tmp = -p1 / 2.0;
    sig1 = tmp + disc;
    sig2 = tmp - disc;
    if (std::abs(sig1) < std::abs(sig2)) std::swap(sig1, sig2);
    auto vec = [&](double sig) {
        const double c = sig * sig + M(0, 0) * M(1, 1) - sig * (M(0, 0) + M(1, 1)) - M(0, 1) * M(0, 1);
        const double a1 = (sig * M(0, 2) + M(0, 1) * M(1, 2) - M(0, 2) * M(1, 1)) / c;
        const double a2 = (sig * M(1, 2) + M(0, 1) * M(0, 2) - M(0, 0) * M(1, 2)) / c;
        const double n = 1.0 / std::sqrt(1 + a1 * a1 + a2 * a2);
        return mk3(a1 * n, a2 * n, n);
    };
    e1 = vec(sig1);
    e2 = vec(sig2);
}
// sum over the entries (column-major) of the element-wise product
double array_prod_sum(const Mat3 &A, const Mat3 &B) {
    double s = 0.0;
    bool first = true;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            const double v = A(r, c) * B(r, c);
            s = first ? v : s + v;
            first = false;
        }
    return s;
}
// matrix of cofactor columns: [c1 x c2, c2 x c0, c0 x c1]  (:95-96)
Mat3 cross_columns(const Mat3 &D) {
    Mat3 R;
    set_col(R, 0, cross(col(D, 1), col(D, 2)));
    set_col(R, 1, cross(col(D, 2), col(D, 0)));
    set_col(R, 2, cross(col(D, 0), col(D, 1)));
    return R;
}
} // namespace

// p3p_lambdatwist.cc:68-244.  lambda_i x_i = R X_i + t with positive lambda_i; x: unit bearings.
int p3p_lambdatwist(const std::vector<Vec3> &x, const std::vector<Vec3> &X, std::vector<CameraPose> *output) {
    const Vec3 dX12 = X[0] - X[1], dX13 = X[0] - X[2], dX23 = X[1] - X[2];
    const double a12 = sqnorm(dX12), b12 = dot(x[0], x[1]);
    const double a13 = sqnorm(dX13), b13 = dot(x[0], x[2]);
    const double a23 = sqnorm(dX23), b23 = dot(x[1], x[2]);
    const double a23b12 = a23 * b12, a12b23 = a12 * b23, a23b13 = a23 * b13, a13b23 = a13 * b23;
    Mat3 D1, D2; // :89-92
    D1(0, 0) = a23;     D1(0, 1) = -a23b12;   D1(0, 2) = 0.0;
    D1(1, 0) = -a23b12; D1(1, 1) = a23 - a12; D1(1, 2) = a12b23;
    D1(2, 0) = 0.0;     D1(2, 1) = a12b23;    D1(2, 2) = -a12;
    D2(0, 0) = a23;     D2(0, 1) = 0.0;       D2(0, 2) = -a23b13;
    D2(1, 0) = 0.0;     D2(1, 1) = -a13;      D2(1, 2) = a13b23;
    D2(2, 0) = -a23b13; D2(2, 1) = a13b23;    D2(2, 2) = a23 - a13;
    const Mat3 DX1 = cross_columns(D1), DX2 = cross_columns(D2);
    // p(gamma) = det(D1 + gamma D2)  (:98-103)
    const double c3 = dot(col(D2, 0), col(DX2, 0));
    double c2 = array_prod_sum(D1, DX2);
    double c1 = array_prod_sum(D2, DX1);
    double c0 = dot(col(D1, 0), col(DX1, 0));
    const double c3inv = 1.0 / c3;
    c2 *= c3inv;
    c1 *= c3inv;
    c0 *= c3inv;
    // one real root in closed form (:110-121), one Newton step (:123-126)
    double a = c1 - c2 * c2 / 3.0;
    double b = (2.0 * c2 * c2 * c2 - 9.0 * c2 * c1) / 27.0 + c0;
    double c = b * b / 4.0 + a * a * a / 27.0;
    double gamma;
    if (c > 0) {
        c = std::sqrt(c);
        b *= -0.5;
        gamma = std::cbrt(b + c) + std::cbrt(b - c) - c2 / 3.0;
    } else {
        c = 3.0 * b / (2.0 * a) * std::sqrt(-3.0 / a);
        gamma = 2.0 * std::sqrt(-a / 3.0) * std::cos(std::acos(c) / 3.0) - c2 / 3.0;
    }
    const double f = gamma * gamma * gamma + c2 * gamma * gamma + c1 * gamma + c0;
    const double df = 3.0 * gamma * gamma + 2.0 * c2 * gamma + c1;
    gamma = gamma - f / df;

    Mat3 D0;
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) D0(r, cc) = D1(r, cc) + gamma * D2(r, cc);
    Vec3 e1, e2;
    double sig1, sig2;
    eig3x3_known0(D0, e1, e2, sig1, sig2);
    double s = std::sqrt(-sig2 / sig1);

    output->clear();
    Mat3 XX;
    set_col(XX, 0, dX12);
    set_col(XX, 1, dX13);
    set_col(XX, 2, cross(dX12, dX13));
    XX = inverse3(XX);
    const double TOL_DOUBLE_ROOT = 1e-12;
    auto emit = [&](double l1, double l2, double l3) {
        refine_lambda(l1, l2, l3, a12, a13, a23, b12, b13, b23);
        const Vec3 v1 = l1 * x[0] - l2 * x[1];
        const Vec3 v2 = l1 * x[0] - l3 * x[2];
        Mat3 YY;
        set_col(YY, 0, v1);
        set_col(YY, 1, v2);
        set_col(YY, 2, cross(v1, v2));
        const Mat3 R = YY * XX;
        output->push_back(pose_from_Rt(R, l1 * x[0] - R * X[0]));
    };
    for (int s_flip = 0; s_flip < 2; ++s_flip, s = -s) {
        // [u1 u2 u3] [lambda1; lambda2; lambda3] = 0  (:153-155)
        const double u1 = e1[0] - s * e2[0], u2 = e1[1] - s * e2[1], u3 = e1[2] - s * e2[2];
        const bool switch_12 = std::abs(u1) < std::abs(u2);
        double qa, qb, qc, w0, w1;
        if (switch_12) { // solve for lambda2 (:164-205)
            w0 = -u1 / u2;
            w1 = -u3 / u2;
            qa = -a13 * w1 * w1 + 2 * a13b23 * w1 - a13 + a23;
            qb = 2 * a13b23 * w0 - 2 * a23b13 - 2 * a13 * w0 * w1;
            qc = -a13 * w0 * w0 + a23;
        } else { // lambda1 as a combination of lambda2 and lambda3 (:207-236)
            w0 = -u2 / u1;
            w1 = -u3 / u1;
            qa = (a13 - a12) * w1 * w1 + 2.0 * a12 * b13 * w1 - a12;
            qb = -2.0 * a13 * b12 * w1 + 2.0 * a12 * b13 * w0 - 2.0 * w0 * w1 * (a12 - a13);
            qc = (a13 - a12) * w0 * w0 - 2.0 * a13 * b12 * w0 + a13;
        }
        const double b2m4ac = qb * qb - 4.0 * qa * qc;
        if (b2m4ac < -TOL_DOUBLE_ROOT) continue; // slightly negative discriminants count as a double root
        const double sq = std::sqrt(std::max(0.0, b2m4ac));
        double tau = (qb > 0) ? (2.0 * qc) / (-qb - sq) : (2.0 * qc) / (-qb + sq);
        for (int tau_flip = 0; tau_flip < 2; ++tau_flip, tau = qc / (qa * tau)) {
            if (tau > 0) {
                if (switch_12) {
                    const double l1 = std::sqrt(a13 / (tau * (tau - 2.0 * b13) + 1.0));
                    const double l3 = tau * l1;
                    const double l2 = w0 * l1 + w1 * l3;
                    if (l2 < 0) continue; // (the reference's `continue` still evaluates the loop increment)
                    emit(l1, l2, l3);
                } else {
                    const double l2 = std::sqrt(a23 / (tau * (tau - 2.0 * b23) + 1.0));
                    const double l3 = tau * l2;
                    const double l1 = w0 * l2 + w1 * l3;
                    if (l1 < 0) continue;
                    emit(l1, l2, l3);
                }
            }
            if (b2m4ac < TOL_DOUBLE_ROOT) break; // double root: the second tau is the same
        }
    }
    return (int)output->size();
}

// ============================ solvers/relpose_5pt.cc ==========================================
namespace {
// Monomial bookkeeping for polynomials in (x,y,z) with homogenising slot 3 ("1").
// Linear  : [x, y, z, 1]
// Quadratic order (relpose_5pt.cc:11-12): [x^2, xy, xz, x, y^2, yz, y, z^2, z, 1]
// Cubic order (relpose_5pt.cc:54-55, Nister): [x^3, y^3, x^2y, xy^2, x^2z, x^2, y^2z, y^2, xyz, xy,
//                                               xz^2, xz, x, yz^2, yz, y, z^3, z^2, z, 1]
struct MonoTables {
    int quad[4][4];  // index of lin_i * lin_j in the quadratic basis
    int cub[10][4];  // index of quad_q * lin_l in the cubic basis
    MonoTables() {
        const int qexp[10][3] = {{2, 0, 0}, {1, 1, 0}, {1, 0, 1}, {1, 0, 0}, {0, 2, 0},
                                 {0, 1, 1}, {0, 1, 0}, {0, 0, 2}, {0, 0, 1}, {0, 0, 0}};
        const int cexp[20][3] = {{3, 0, 0}, {0, 3, 0}, {2, 1, 0}, {1, 2, 0}, {2, 0, 1}, {2, 0, 0}, {0, 2, 1},
                                 {0, 2, 0}, {1, 1, 1}, {1, 1, 0}, {1, 0, 2}, {1, 0, 1}, {1, 0, 0}, {0, 1, 2},
                                 {0, 1, 1}, {0, 1, 0}, {0, 0, 3}, {0, 0, 2}, {0, 0, 1}, {0, 0, 0}};
        const int lexp[4][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
        auto find = [](const int (*tab)[3], int n, int a, int b, int c) {
            for (int i = 0; i < n; ++i)
                if (tab[i][0] == a && tab[i][1] == b && tab[i][2] == c) return i;
            return -1;
        };
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                quad[i][j] = find(qexp, 10, lexp[i][0] + lexp[j][0], lexp[i][1] + lexp[j][1], lexp[i][2] + lexp[j][2]);
        for (int q = 0; q < 10; ++q)
            for (int l = 0; l < 4; ++l)
                cub[q][l] = find(cexp, 20, qexp[q][0] + lexp[l][0], qexp[q][1] + lexp[l][1], qexp[q][2] + lexp[l][2]);
    }
};
const MonoTables MT;

// ---- optional "reference operation order" (test hook, off by default) ------------------------------------------
// By default the 5- and 7-point solvers restate the reference's GENERATED polynomial expansions structurally (the code
// below), which is algebraically identical but adds the terms up in a different order.  tests/test_ref_sources.py
// switches this mode on to prove that the order is the ONLY difference to the reference's sources:
//   * o1 / o1p / o1m / o2 / o2p (relpose_5pt.cc:13-99): every output monomial is the sum of its products in ascending
//     index of the first factor, formed FIRST and then assigned / added / subtracted (a rule, implemented below);
//   * the cubic of relpose_7pt.cc:22-37: terms in lexicographic order of (entry, basis) triples (a rule, below);
//   * the degree-10 determinant of relpose_5pt.cc:191-352: the term order of that computer-algebra output follows no
//     simple rule; the test parses it from the reference's source file at run time and injects it
//     (plo_set_reference_order) — nothing of it is stored in this repository.  The injected table is checked against the
//     mathematically complete term set before it is accepted.
struct DetTerm { int sign, r[3], c[3]; };
struct RefOrder {
    bool enabled = false;
    std::vector<DetTerm> det[11];
};
RefOrder g_ref_order;

struct OrderedTables { // products contributing to each output monomial, ascending in the first factor's index
    std::vector<std::pair<int, int>> quad[10], cub[20];
    OrderedTables() {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) quad[MT.quad[i][j]].push_back({i, j});
        for (int q = 0; q < 10; ++q)
            for (int l = 0; l < 4; ++l) cub[MT.cub[q][l]].push_back({q, l});
    }
};
const OrderedTables OT;
// mode: 0 assign, +1 add, -1 subtract the per-monomial sum
inline void lin_mul_ref(const double a[4], const double b[4], double c[10], int mode) {
    for (int m = 0; m < 10; ++m) {
        const auto &t = OT.quad[m];
        double s = a[t[0].first] * b[t[0].second];
        for (size_t k = 1; k < t.size(); ++k) s = s + a[t[k].first] * b[t[k].second];
        c[m] = mode == 0 ? s : (mode > 0 ? c[m] + s : c[m] - s);
    }
}
inline void quad_mul_ref(const double a[10], const double b[4], double c[20], int mode) {
    for (int m = 0; m < 20; ++m) {
        const auto &t = OT.cub[m];
        double s = a[t[0].first] * b[t[0].second];
        for (size_t k = 1; k < t.size(); ++k) s = s + a[t[k].first] * b[t[k].second];
        c[m] = mode == 0 ? s : c[m] + s;
    }
}

// c (+)= sgn * a*b for linear a,b -> quadratic c      (o1/o1p/o1m, relpose_5pt.cc:13-48)
inline void lin_mul_acc(const double a[4], const double b[4], double c[10], double sgn) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) c[MT.quad[i][j]] += sgn * (a[i] * b[j]);
}
// c += a*b for quadratic a, linear b -> cubic c         (o2/o2p, relpose_5pt.cc:56-99)
inline void quad_mul_acc(const double a[10], const double b[4], double c[20]) {
    for (int q = 0; q < 10; ++q)
        for (int l = 0; l < 4; ++l) c[MT.cub[q][l]] += a[q] * b[l];
}

// relpose_5pt.cc:101-157.  Nb[k][r] = coefficient of basis matrix r (x,y,z,1) in entry k (col-major) of E.
void compute_trace_constraints_ref_order(const double Nb[9][4], double coeffs[10][20]) {
    auto EE = [&](int i, int j) -> const double * { return Nb[3 * j + i]; };
    double d[10];
    double *row = coeffs[9]; // determinant constraint (:113-125)
    lin_mul_ref(EE(0, 1), EE(1, 2), d, 0);
    lin_mul_ref(EE(0, 2), EE(1, 1), d, -1);
    quad_mul_ref(d, EE(2, 0), row, 0);
    lin_mul_ref(EE(0, 2), EE(1, 0), d, 0);
    lin_mul_ref(EE(0, 0), EE(1, 2), d, -1);
    quad_mul_ref(d, EE(2, 1), row, 1);
    lin_mul_ref(EE(0, 0), EE(1, 1), d, 0);
    lin_mul_ref(EE(0, 1), EE(1, 0), d, -1);
    quad_mul_ref(d, EE(2, 2), row, 1);
    double EET[3][3][10]; // (:129-136)
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            lin_mul_ref(EE(i, 0), EE(j, 0), EET[i][j], 0);
            lin_mul_ref(EE(i, 1), EE(j, 1), EET[i][j], 1);
            lin_mul_ref(EE(i, 2), EE(j, 2), EET[i][j], 1);
        }
    for (int m = 0; m < 10; ++m) { // (:139-144)
        const double t = 0.5 * (EET[0][0][m] + EET[1][1][m] + EET[2][2][m]);
        EET[0][0][m] -= t;
        EET[1][1][m] -= t;
        EET[2][2][m] -= t;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < i; ++j) std::copy(EET[j][i], EET[j][i] + 10, EET[i][j]);
    int cnt = 0; // (:146-154)
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double *r = coeffs[cnt++];
            quad_mul_ref(EET[i][0], EE(0, j), r, 0);
            quad_mul_ref(EET[i][1], EE(1, j), r, 1);
            quad_mul_ref(EET[i][2], EE(2, j), r, 1);
        }
}
void compute_trace_constraints(const double Nb[9][4], double coeffs[10][20]) {
    if (g_ref_order.enabled) return compute_trace_constraints_ref_order(Nb, coeffs);
    auto EE = [&](int i, int j) -> const double * { return Nb[3 * j + i]; };
    // determinant constraint -> row 9 (:113-125): cofactor expansion along the last row of E
    {
        double *row = coeffs[9];
        std::fill(row, row + 20, 0.0);
        double d[10];
        std::fill(d, d + 10, 0.0);
        lin_mul_acc(EE(0, 1), EE(1, 2), d, 1.0);
        lin_mul_acc(EE(0, 2), EE(1, 1), d, -1.0);
        quad_mul_acc(d, EE(2, 0), row);
        std::fill(d, d + 10, 0.0);
        lin_mul_acc(EE(0, 2), EE(1, 0), d, 1.0);
        lin_mul_acc(EE(0, 0), EE(1, 2), d, -1.0);
        quad_mul_acc(d, EE(2, 1), row);
        std::fill(d, d + 10, 0.0);
        lin_mul_acc(EE(0, 0), EE(1, 1), d, 1.0);
        lin_mul_acc(EE(0, 1), EE(1, 0), d, -1.0);
        quad_mul_acc(d, EE(2, 2), row);
    }
    // EE^T (symmetric, quadratic entries) (:129-136) and trace subtraction (:139-144)
    double EET[3][3][10];
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            std::fill(EET[i][j], EET[i][j] + 10, 0.0);
            for (int k = 0; k < 3; ++k) lin_mul_acc(EE(i, k), EE(j, k), EET[i][j], 1.0);
        }
    for (int m = 0; m < 10; ++m) {
        const double t = 0.5 * (EET[0][0][m] + EET[1][1][m] + EET[2][2][m]);
        EET[0][0][m] -= t;
        EET[1][1][m] -= t;
        EET[2][2][m] -= t;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < i; ++j) std::copy(EET[j][i], EET[j][i] + 10, EET[i][j]);
    // (EE^T - 1/2 tr) E = 0 -> rows 0..8 (:146-154)
    int cnt = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double *row = coeffs[cnt++];
            std::fill(row, row + 20, 0.0);
            for (int k = 0; k < 3; ++k) quad_mul_acc(EET[i][k], EE(k, j), row);
        }
}

// ascending-power polynomial helpers for the 3x3 determinant (relpose_5pt.cc:191-352)
inline void pmul(const double *a, int da, const double *b, int db, double *c) {
    for (int i = 0; i <= da + db; ++i) c[i] = 0.0;
    for (int i = 0; i <= da; ++i)
        for (int j = 0; j <= db; ++j) c[i + j] += a[i] * b[j];
}
} // namespace

// Test hook: when set, relpose_5pt copies its intermediate results there — Nb (36: Nb[4k + r]) | A (39, row-major
// 3 x 13) | determinant polynomial (11, ascending).  tests/test_solver5_lane_host.py pins the device solver's
// thread-per-sample first half (poselib_b200/csrc/solver5_lane.cuh, host build) to these bit for bit.
thread_local double *g_relpose_5pt_stage_out = nullptr;

// relpose_5pt.cc:159-395
int relpose_5pt(const std::vector<Vec3> &x1, const std::vector<Vec3> &x2, std::vector<Mat3> *essential_matrices) {
    // 9x5 epipolar constraint matrix, column-major (:163-166)
    double M[9 * 5];
    for (int i = 0; i < 5; ++i)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) M[i * 9 + 3 * a + b] = x1[i][a] * x2[i][b];
    double Q[81];
    full_piv_householder_q<9, 5>(M, Q); // :167
    double Nb[9][4];                    // N = Q.rightCols(4)^T  (:168): Nb[k][r] = Q(k, 5+r)
    for (int k = 0; k < 9; ++k)
        for (int r = 0; r < 4; ++r) Nb[k][r] = Q[(5 + r) * 9 + k];

    double coeffs[10][20];
    compute_trace_constraints(Nb, coeffs); // :171-172
    {                                      // :173  coeffs[:,10:] = lu(coeffs[:,:10]).solve(coeffs[:,10:])
        double Aleft[10 * 10], Bright[10 * 10];
        for (int r = 0; r < 10; ++r)
            for (int c = 0; c < 10; ++c) {
                Aleft[r * 10 + c] = coeffs[r][c];
                Bright[r * 10 + c] = coeffs[r][10 + c];
            }
        partial_piv_lu_solve(10, Aleft, 10, Bright, 10, 10);
        for (int r = 0; r < 10; ++r)
            for (int c = 0; c < 10; ++c) coeffs[r][10 + c] = Bright[r * 10 + c];
    }
    // eliminations using the 6 bottom rows (:176-189)
    double A[3][13];
    for (int i = 0; i < 3; ++i) {
        const double *top = coeffs[4 + 2 * i] + 10, *bot = coeffs[5 + 2 * i] + 10;
        A[i][0] = 0.0;
        for (int k = 0; k < 3; ++k) A[i][1 + k] = top[k];
        for (int k = 0; k < 3; ++k) A[i][0 + k] -= bot[k];
        A[i][4] = 0.0;
        for (int k = 0; k < 3; ++k) A[i][5 + k] = top[3 + k];
        for (int k = 0; k < 3; ++k) A[i][4 + k] -= bot[3 + k];
        A[i][8] = 0.0;
        for (int k = 0; k < 4; ++k) A[i][9 + k] = top[6 + k];
        for (int k = 0; k < 4; ++k) A[i][8 + k] -= bot[6 + k];
    }
    // degree-10 determinant polynomial c[0..10], ascending (:191-352).  Row i of A holds
    // p_i0 (deg 3: A[i][0..3]), p_i1 (deg 3: A[i][4..7]), p_i2 (deg 4: A[i][8..12]), highest power first.
    double p[3][3][5];
    for (int i = 0; i < 3; ++i) {
        for (int k = 0; k <= 3; ++k) p[i][0][k] = A[i][3 - k];
        for (int k = 0; k <= 3; ++k) p[i][1][k] = A[i][7 - k];
        for (int k = 0; k <= 4; ++k) p[i][2][k] = A[i][12 - k];
    }
    double c[11];
    if (g_ref_order.enabled) { // injected term order of the reference's expansion (see RefOrder above)
        for (int k = 0; k <= 10; ++k) {
            double acc = 0.0;
            bool first = true;
            for (const DetTerm &t : g_ref_order.det[k]) {
                const double v = A[t.r[0]][t.c[0]] * A[t.r[1]][t.c[1]] * A[t.r[2]][t.c[2]];
                if (first) acc = t.sign > 0 ? v : -v;
                else acc = t.sign > 0 ? acc + v : acc - v;
                first = false;
            }
            c[k] = acc;
        }
    } else {
        double m1[8], m2[8], minor[8], term[11];
        for (int k = 0; k <= 10; ++k) c[k] = 0.0;
        // + p00 * (p11*p22 - p12*p21)
        pmul(p[1][1], 3, p[2][2], 4, m1);
        pmul(p[1][2], 4, p[2][1], 3, m2);
        for (int k = 0; k <= 7; ++k) minor[k] = m1[k] - m2[k];
        pmul(p[0][0], 3, minor, 7, term);
        for (int k = 0; k <= 10; ++k) c[k] += term[k];
        // - p01 * (p10*p22 - p12*p20)
        pmul(p[1][0], 3, p[2][2], 4, m1);
        pmul(p[1][2], 4, p[2][0], 3, m2);
        for (int k = 0; k <= 7; ++k) minor[k] = m1[k] - m2[k];
        pmul(p[0][1], 3, minor, 7, term);
        for (int k = 0; k <= 10; ++k) c[k] -= term[k];
        // + p02 * (p10*p21 - p11*p20)
        double m3[7], m4[7], minor2[7];
        pmul(p[1][0], 3, p[2][1], 3, m3);
        pmul(p[1][1], 3, p[2][0], 3, m4);
        for (int k = 0; k <= 6; ++k) minor2[k] = m3[k] - m4[k];
        pmul(p[0][2], 4, minor2, 6, term);
        for (int k = 0; k <= 10; ++k) c[k] += term[k];
    }
    if (g_relpose_5pt_stage_out) {
        double *o = g_relpose_5pt_stage_out;
        for (int k = 0; k < 9; ++k)
            for (int r = 0; r < 4; ++r) o[4 * k + r] = Nb[k][r];
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 13; ++k) o[36 + 13 * i + k] = A[i][k];
        for (int k = 0; k <= 10; ++k) o[75 + k] = c[k];
    }
    double roots[10];
    const int n_sols = bisect_sturm10(c, roots); // :356

    // back-substitution (:359-392)
    for (int i = 0; i < n_sols; ++i) {
        const double z = roots[i], z2 = z * z, z3 = z2 * z, z4 = z2 * z2;
        double B[3][2], bb[3];
        for (int r = 0; r < 3; ++r) {
            B[r][0] = A[r][0] * z3 + A[r][1] * z2 + A[r][2] * z + A[r][3];
            B[r][1] = A[r][4] * z3 + A[r][5] * z2 + A[r][6] * z + A[r][7];
            bb[r] = A[r][8] * z4 + A[r][9] * z3 + A[r][10] * z2 + A[r][11] * z + A[r][12];
        }
        // 2x2 inverse of the top rows (:377)  [Eigen: adjugate / det]
        const double det = B[0][0] * B[1][1] - B[0][1] * B[1][0];
        const double invdet = 1.0 / det;
        const double i00 = B[1][1] * invdet, i01 = -B[0][1] * invdet, i10 = -B[1][0] * invdet, i11 = B[0][0] * invdet;
        double xz[2] = {i00 * bb[0] + i01 * bb[1], i10 * bb[0] + i11 * bb[1]};
        if (std::abs(B[2][0] * xz[0] + B[2][1] * xz[1] - bb[2]) > 1e-6) col_piv_qr_solve_3x2(B, bb, xz); // :380-382
        const double x = -xz[0], y = -xz[1];
        const double inv_norm = 1.0 / std::sqrt(x * x + y * y + z * z + 1.0); // :388
        Mat3 E;
        for (int k = 0; k < 9; ++k) {
            const double e = Nb[k][0] * x + Nb[k][1] * y + Nb[k][2] * z + Nb[k][3]; // :385
            E(k % 3, k / 3) = e * inv_norm;                                          // column-major map (:363)
        }
        essential_matrices->push_back(E);
    }
    return n_sols;
}
// relpose_5pt.cc:397-409
int relpose_5pt(const std::vector<Vec3> &x1, const std::vector<Vec3> &x2, std::vector<CameraPose> *output) {
    std::vector<Mat3> Es;
    const int n_sols = relpose_5pt(x1, x2, &Es);
    output->clear();
    for (int i = 0; i < n_sols; ++i) motion_from_essential(Es[i], x1, x2, output);
    return (int)output->size();
}

// ============================ solvers/relpose_7pt.cc ==========================================
// relpose_7pt.cc:10-60
int relpose_7pt(const std::vector<Vec3> &x1, const std::vector<Vec3> &x2, std::vector<Mat3> *Fs) {
    double M[9 * 7];
    for (int i = 0; i < 7; ++i)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) M[i * 9 + 3 * a + b] = x1[i][a] * x2[i][b];
    double Q[81];
    full_piv_householder_q<9, 7>(M, Q);
    // N = Q.rightCols(2) (:19); F(x) = x*F0 + F1, entries column-major
    const double *n0 = Q + 7 * 9, *n1 = Q + 8 * 9;
    // det(x F0 + F1) = c3 x^3 + c2 x^2 + c1 x + c0 via mixed determinants (:22-37)
    auto det_cols = [&](const double *a, const double *b, const double *c) {
        // determinant of the 3x3 matrix whose column-major 9-vector takes column j from {a,b,c}[j]
        const double *col0 = a, *col1 = b + 3, *col2 = c + 6;
        return col0[0] * (col1[1] * col2[2] - col1[2] * col2[1]) - col1[0] * (col0[1] * col2[2] - col0[2] * col2[1]) +
               col2[0] * (col0[1] * col1[2] - col0[2] * col1[1]);
    };
    double c3 = det_cols(n0, n0, n0);
    double c2 = det_cols(n1, n0, n0) + det_cols(n0, n1, n0) + det_cols(n0, n0, n1);
    double c1 = det_cols(n0, n1, n1) + det_cols(n1, n0, n1) + det_cols(n1, n1, n0);
    double c0 = det_cols(n1, n1, n1);
    if (g_ref_order.enabled) {
        // relpose_7pt.cc:22-37 as a rule: the six Leibniz products of det(x F0 + F1) over column-major entries
        // (0,4,8)+ (0,5,7)- (1,3,8)- (1,5,6)+ (2,3,7)+ (2,4,6)-, each expanded over the basis choice (0 = F0, 1 = F1) of its
        // three factors; the coefficient of x^(3-p) collects the expansions with p factors from F1, ordered
        // lexicographically by (entry, basis, entry, basis, entry, basis), products and sums left to right.
        static const int T[6][4] = {{0, 4, 8, +1}, {0, 5, 7, -1}, {1, 3, 8, -1}, {1, 5, 6, +1}, {2, 3, 7, +1}, {2, 4, 6, -1}};
        struct Tm { int key[6]; int sign; };
        double cc[4];
        for (int pcount = 0; pcount <= 3; ++pcount) {
            std::vector<Tm> terms;
            for (int t = 0; t < 6; ++t)
                for (int mask = 0; mask < 8; ++mask) {
                    const int b0 = (mask >> 2) & 1, b1 = (mask >> 1) & 1, b2 = mask & 1;
                    if (b0 + b1 + b2 != pcount) continue;
                    terms.push_back({{T[t][0], b0, T[t][1], b1, T[t][2], b2}, T[t][3]});
                }
            std::sort(terms.begin(), terms.end(), [](const Tm &a, const Tm &b) {
                return std::lexicographical_compare(a.key, a.key + 6, b.key, b.key + 6);
            });
            double acc = 0.0;
            bool first = true;
            for (const Tm &t : terms) {
                auto N = [&](int e, int b) { return b ? n1[e] : n0[e]; };
                const double v = N(t.key[0], t.key[1]) * N(t.key[2], t.key[3]) * N(t.key[4], t.key[5]);
                if (first) acc = t.sign > 0 ? v : -v;
                else acc = t.sign > 0 ? acc + v : acc - v;
                first = false;
            }
            cc[pcount] = acc;
        }
        c3 = cc[0]; c2 = cc[1]; c1 = cc[2]; c0 = cc[3];
    }
    double roots[3];
    int n_roots;
    if (std::abs(c3) < 1e-14) { // :42-44
        n_roots = solve_quadratic_real(c2, c1, c0, roots);
    } else {
        const double inv_c3 = 1.0 / c3;
        n_roots = solve_cubic_real(c2 * inv_c3, c1 * inv_c3, c0 * inv_c3, roots);
    }
    Fs->clear();
    for (int i = 0; i < n_roots; ++i) { // :53-57
        double f[9], n2 = 0;
        for (int k = 0; k < 9; ++k) {
            f[k] = n0[k] * roots[i] + n1[k];
            n2 += f[k] * f[k];
        }
        if (n2 > 0) {
            const double n = std::sqrt(n2);
            for (int k = 0; k < 9; ++k) f[k] /= n;
        }
        Mat3 F;
        for (int k = 0; k < 9; ++k) F(k % 3, k / 3) = f[k];
        Fs->push_back(F);
    }
    return n_roots;
}

// ============================ solvers/relpose_8pt.cc ==========================================
// relpose_8pt.cc:52-83.  Row i of the n x 9 system is [x2.x * x1^T, x2.y * x1^T, x2.z * x1^T] (:41-49), i.e. the
// ROW-major entries of E.  Exactly 8 points: last column of the Householder Q of the transposed system (:63-67);
// more: eigenvector of the smallest eigenvalue of A^T A (:68-72, an iterative solver in Eigen -> tolerance parity).
// Then the closest essential matrix in Frobenius norm: singular values (a, b, c) -> ((a+b)/2, (a+b)/2, 0) (:74-81).
void essential_matrix_8pt(const std::vector<Vec3> &x1, const std::vector<Vec3> &x2, Mat3 *essential_matrix) {
    const size_t n = x1.size();
    double e[9];
    auto row = [&](size_t i, double r[9]) {
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) r[3 * a + b] = x2[i][a] * x1[i][b];
    };
    if (n == 8) {
        double At[9 * 8], Q[81]; // transposed system, 9 x 8 column-major: column i = row i of A
        for (size_t i = 0; i < 8; ++i) row(i, At + 9 * i);
        householder_q<9, 8>(At, Q);
        for (int k = 0; k < 9; ++k) e[k] = Q[8 * 9 + k];
    } else {
        double G[81]; // A^T A, row-major; entry (p,q) summed over the correspondences in order
        std::vector<double> rows(9 * n);
        for (size_t i = 0; i < n; ++i) row(i, rows.data() + 9 * i);
        for (int p = 0; p < 9; ++p)
            for (int q = 0; q < 9; ++q) {
                double s = 0.0;
                for (size_t i = 0; i < n; ++i) s = (i == 0) ? rows[9 * i + p] * rows[9 * i + q] : s + rows[9 * i + p] * rows[9 * i + q];
                G[p * 9 + q] = s;
            }
        double evals[9], V[81];
        sym_eigen_jacobi<9>(G, evals, V);
        for (int k = 0; k < 9; ++k) e[k] = V[k]; // first column = smallest eigenvalue
    }
    Mat3 E;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) E(a, b) = e[3 * a + b]; // Map<const RowMajor 3x3>
    Mat3 U, V;
    double d[3];
    svd3(E, U, d, V);
    const double m = (d[0] + d[1]) / 2.;
    Mat3 UD; // U * diag(m, m, 0)
    for (int r = 0; r < 3; ++r) {
        UD(r, 0) = U(r, 0) * m;
        UD(r, 1) = U(r, 1) * m;
        UD(r, 2) = U(r, 2) * 0.0;
    }
    *essential_matrix = UD * transpose(V);
}
// relpose_8pt.cc:85-94
int relpose_8pt(const std::vector<Vec3> &x1, const std::vector<Vec3> &x2, std::vector<CameraPose> *output) {
    Mat3 E;
    essential_matrix_8pt(x1, x2, &E);
    output->clear();
    motion_from_essential(E, x1, x2, output);
    return (int)output->size();
}

// ============================ solvers/homography_4pt.cc =======================================
// homography_4pt.cc:36-128 (SKS/ACA closed form, Cai et al. PAMI'25)
int homography_4pt(const std::vector<Vec3> &x1, const std::vector<Vec3> &x2, Mat3 *H, bool check_cheir) {
    if (check_cheir) { // :38-55
        Vec3 p = cross(x1[0], x1[1]), q = cross(x2[0], x2[1]);
        if (dot(p, x1[2]) * dot(q, x2[2]) < 0) return 0;
        if (dot(p, x1[3]) * dot(q, x2[3]) < 0) return 0;
        p = cross(x1[2], x1[3]);
        q = cross(x2[2], x2[3]);
        if (dot(p, x1[0]) * dot(q, x2[0]) < 0) return 0;
        if (dot(p, x1[1]) * dot(q, x2[1]) < 0) return 0;
    }
    double ax[4], ay[4], bx[4], by[4];
    for (int i = 0; i < 4; ++i) { // :59-62
        ax[i] = x1[i][0] / x1[i][2];
        ay[i] = x1[i][1] / x1[i][2];
        bx[i] = x2[i][0] / x2[i][2];
        by[i] = x2[i][1] / x2[i][2];
    }
    // source plane: edge vectors from point 0 and the affine factor (:64-68)
    const double sNx = ax[1] - ax[0], sPx = ax[2] - ax[0], sQx = ax[3] - ax[0];
    const double sNy = ay[1] - ay[0], sPy = ay[2] - ay[0], sQy = ay[3] - ay[0];
    const double fA1 = sNx * sPy - sNy * sPx;
    const double Q3x = sPy * sQx - sPx * sQy;
    const double Q3y = sNx * sQy - sNy * sQx;
    // target plane (:75-79)
    const double tNx = bx[1] - bx[0], tPx = bx[2] - bx[0], tQx = bx[3] - bx[0];
    const double tNy = by[1] - by[0], tPy = by[2] - by[0], tQy = by[3] - by[0];
    const double fA2 = tNx * tPy - tNy * tPx;
    const double Q4x = tPy * tQx - tPx * tQy;
    const double Q4y = tNx * tQy - tNy * tQx;
    // core transformation (:85-90)
    const double tt1 = fA1 - Q3x - Q3y;
    const double C11 = Q3y * Q4x * tt1;
    const double C22 = Q3x * Q4y * tt1;
    const double C33 = Q3x * Q3y * (fA2 - Q4x - Q4y);
    const double C31 = C11 - C33, C32 = C22 - C33;
    // H_A2^{-1} H_C upper-left block (:96-101)
    const double tt3 = bx[0] * C33, tt4 = by[0] * C33;
    const double H1_11 = bx[1] * C11 - tt3, H1_12 = bx[2] * C22 - tt3;
    const double H1_21 = by[1] * C11 - tt4, H1_22 = by[2] * C22 - tt4;
    double h[9]; // row-major H (:110-120)
    h[0] = H1_11 * sPy - H1_12 * sNy;
    h[1] = H1_12 * sNx - H1_11 * sPx;
    h[3] = H1_21 * sPy - H1_22 * sNy;
    h[4] = H1_22 * sNx - H1_21 * sPx;
    h[6] = C31 * sPy - C32 * sNy;
    h[7] = C32 * sNx - C31 * sPx;
    h[2] = tt3 * fA1 - h[0] * ax[0] - h[1] * ay[0];
    h[5] = tt4 * fA1 - h[3] * ax[0] - h[4] * ay[0];
    h[8] = C33 * fA1 - h[6] * ax[0] - h[7] * ay[0];
    Mat3 Hm;
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) Hm(r, cc) = h[3 * r + cc];
    // H->normalize() : Frobenius, column-major accumulation (:121)
    double n2 = 0;
    for (int cc = 0; cc < 3; ++cc)
        for (int r = 0; r < 3; ++r) n2 += Hm(r, cc) * Hm(r, cc);
    if (n2 > 0) {
        const double n = std::sqrt(n2);
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc) Hm(r, cc) /= n;
    }
    *H = Hm;
    if (std::abs(det3(Hm)) < 1e-8) return 0; // :122-125
    return 1;
}

} // namespace plo

namespace plo {
bool reference_order_enabled() { return g_ref_order.enabled; }
} // namespace plo

// test hooks for the two scalar helpers of p3p (anonymous namespace above), compared with the reference's
// p3p_common.h through oracle/_ref (tests/test_ref_pins.py)
extern "C" {
// Test hook (tests/test_ref_sources.py): switch the 5- / 7-point solvers to the reference's operation order.  `terms`
// holds n rows {k, sign, r0, c0, r1, c1, r2, c2}: coefficient c[k] of the 5-point determinant polynomial is the running
// sum, in row order, of sign * A(r0,c0) * A(r1,c1) * A(r2,c2) (relpose_5pt.cc:191-352, parsed by the test from the
// reference's source at run time).  The table is accepted only if, for every k, it is exactly the complete term set of
// det [p_i0 p_i1 p_i2] (as a signed multiset).  Returns 0 on success; enable = 0 restores the default order.
int plo_set_reference_order(int enable, const int32_t *terms, int n) {
    plo::g_ref_order.enabled = false;
    for (int k = 0; k <= 10; ++k) plo::g_ref_order.det[k].clear();
    if (!enable) return 0;
    auto deg = [](int c) { return c < 4 ? 3 - c : (c < 8 ? 7 - c : 12 - c); };
    auto blk = [](int c) { return c < 4 ? 0 : (c < 8 ? 1 : 2); };
    std::vector<std::array<int, 7>> got[11], want[11]; // canonical: (c of row 0, c of row 1, c of row 2), sign
    for (int t = 0; t < n; ++t) {
        const int32_t *q = terms + 8 * t;
        const int k = q[0], sign = q[1];
        if (k < 0 || k > 10 || (sign != 1 && sign != -1)) return 1;
        plo::DetTerm d;
        d.sign = sign;
        int col_of_row[3] = {-1, -1, -1};
        for (int f = 0; f < 3; ++f) {
            d.r[f] = q[2 + 2 * f];
            d.c[f] = q[3 + 2 * f];
            if (d.r[f] < 0 || d.r[f] > 2 || d.c[f] < 0 || d.c[f] > 12 || col_of_row[d.r[f]] >= 0) return 2;
            col_of_row[d.r[f]] = d.c[f];
        }
        if (deg(d.c[0]) + deg(d.c[1]) + deg(d.c[2]) != k) return 3;
        got[k].push_back({col_of_row[0], col_of_row[1], col_of_row[2], sign, 0, 0, 0});
        plo::g_ref_order.det[k].push_back(d);
    }
    static const int perm[6][4] = {{0, 1, 2, +1}, {0, 2, 1, -1}, {1, 0, 2, -1}, {1, 2, 0, +1}, {2, 0, 1, +1}, {2, 1, 0, -1}};
    for (int pi = 0; pi < 6; ++pi) // row i takes its entry from column block perm[pi][i]
        for (int c0 = 0; c0 < 13; ++c0)
            for (int c1 = 0; c1 < 13; ++c1)
                for (int c2 = 0; c2 < 13; ++c2)
                    if (blk(c0) == perm[pi][0] && blk(c1) == perm[pi][1] && blk(c2) == perm[pi][2])
                        want[deg(c0) + deg(c1) + deg(c2)].push_back({c0, c1, c2, perm[pi][3], 0, 0, 0});
    for (int k = 0; k <= 10; ++k) {
        std::sort(got[k].begin(), got[k].end());
        std::sort(want[k].begin(), want[k].end());
        if (got[k] != want[k]) {
            for (int j = 0; j <= 10; ++j) plo::g_ref_order.det[j].clear();
            return 4;
        }
    }
    plo::g_ref_order.enabled = true;
    return 0;
}
int plo_p3p_root2real(double b, double c, double *r) { return plo::root2real(b, c, r[0], r[1]) ? 1 : 0; }
void plo_p3p_refine_lambda(double *l, double a12, double a13, double a23, double b12, double b13, double b23) {
    plo::refine_lambda(l[0], l[1], l[2], a12, a13, a23, b12, b13, b23);
}
}
