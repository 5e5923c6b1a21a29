// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under poselib_b200/ may include, link or call this.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// PARITY PARTLY PINNED.  The reference (PoseLib @ a69263d) as a whole cannot be compiled here (Eigen3 absent) and its
// tests hold no golden vectors for the solver/scorer arithmetic.  What DOES compile without Eigen
// arithmetic — robust/sampling.cc, the loop templates of robust/ransac_impl.h, misc/univariate.cc, the Sturm templates of
// misc/sturm.h, the scalar helpers of solvers/p3p_common.h and the element-access-only functions of robust/utils.cc — is built from the reference sources into oracle/_ref (oracle/ref/ref_capi.cc) and pins
// the oracle's sampler, loop control flow, iteration arithmetic, cubic/quadratic/p3p scalar solvers, Sturm root isolation,
// F / H scorers, masks and the real-focal check bit for bit (tests/test_ref_pins.py).  The reference's own sources for the
// whole path, compiled unmodified on top of mini-Eigen (oracle/ref/mini: Eigen's interface implemented with THIS
// file's routines), pin the oracle's transcription of PoseLib's logic end to end (oracle/_ref/libplref2.so,
// tests/test_ref_sources.py).  Whether the routines below equal real Eigen's bit for bit stays UNPINNED.  This file
// restates, as plain sequential loops, the small Eigen routines whose arithmetic shapes PoseLib's
// results.  Eigen's source is not available here; semantics are recalled from
// Eigen 3.4 and summation order is canonical left-to-right.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <limits>

namespace plo {

struct Vec2 { double v[2]; double &operator[](int i) { return v[i]; } double operator[](int i) const { return v[i]; } };
struct Vec3 { double v[3]; double &operator[](int i) { return v[i]; } double operator[](int i) const { return v[i]; } };
struct Vec4 { double v[4]; double &operator[](int i) { return v[i]; } double operator[](int i) const { return v[i]; } };
// Row-major 3x3: m[r][c]
struct Mat3 {
    double m[3][3];
    double &operator()(int r, int c) { return m[r][c]; }
    double operator()(int r, int c) const { return m[r][c]; }
};

inline Vec3 mk3(double a, double b, double c) { Vec3 r; r.v[0] = a; r.v[1] = b; r.v[2] = c; return r; }
inline Vec3 operator+(const Vec3 &a, const Vec3 &b) { return mk3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Vec3 operator-(const Vec3 &a, const Vec3 &b) { return mk3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline Vec3 operator-(const Vec3 &a) { return mk3(-a[0], -a[1], -a[2]); }
inline Vec3 operator*(double s, const Vec3 &a) { return mk3(s * a[0], s * a[1], s * a[2]); }
inline Vec3 operator*(const Vec3 &a, double s) { return mk3(a[0] * s, a[1] * s, a[2] * s); }
inline Vec3 operator/(const Vec3 &a, double s) { return mk3(a[0] / s, a[1] / s, a[2] / s); }
inline double dot(const Vec3 &a, const Vec3 &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double sqnorm(const Vec3 &a) { return dot(a, a); }
inline double norm(const Vec3 &a) { return std::sqrt(sqnorm(a)); }
inline Vec3 cross(const Vec3 &a, const Vec3 &b) {
    return mk3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
// Eigen normalized(): divide by sqrt(squaredNorm) when squaredNorm > 0
inline Vec3 normalized(const Vec3 &a) {
    double n2 = sqnorm(a);
    if (n2 > 0) return a / std::sqrt(n2);
    return a;
}
// x.homogeneous().normalized() of a 2D point (estimators/*.cc generate_models)
inline Vec3 bearing(const Vec2 &x) { return normalized(mk3(x[0], x[1], 1.0)); }

inline Mat3 mat3_zero() { Mat3 r; std::memset(&r, 0, sizeof(r)); return r; }
inline Mat3 mat3_identity() { Mat3 r = mat3_zero(); r(0, 0) = r(1, 1) = r(2, 2) = 1.0; return r; }
inline Vec3 col(const Mat3 &A, int c) { return mk3(A(0, c), A(1, c), A(2, c)); }
inline Vec3 row(const Mat3 &A, int r) { return mk3(A(r, 0), A(r, 1), A(r, 2)); }
inline void set_col(Mat3 &A, int c, const Vec3 &v) { A(0, c) = v[0]; A(1, c) = v[1]; A(2, c) = v[2]; }
inline void set_row(Mat3 &A, int r, const Vec3 &v) { A(r, 0) = v[0]; A(r, 1) = v[1]; A(r, 2) = v[2]; }
inline Mat3 operator*(const Mat3 &A, const Mat3 &B) {
    Mat3 C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C(i, j) = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j);
    return C;
}
inline Vec3 operator*(const Mat3 &A, const Vec3 &x) {
    return mk3(A(0, 0) * x[0] + A(0, 1) * x[1] + A(0, 2) * x[2], A(1, 0) * x[0] + A(1, 1) * x[1] + A(1, 2) * x[2],
               A(2, 0) * x[0] + A(2, 1) * x[1] + A(2, 2) * x[2]);
}
inline Mat3 transpose(const Mat3 &A) {
    Mat3 T;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T(i, j) = A(j, i);
    return T;
}
inline Mat3 operator*(const Mat3 &A, double s) {
    Mat3 C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C(i, j) = A(i, j) * s;
    return C;
}
// Matrix3d::determinant(): Laplace along row 0
inline double det3(const Mat3 &m) {
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
           m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
}
// Matrix3d::inverse(): cofactors / determinant
inline Mat3 inverse3(const Mat3 &m) {
    Mat3 c; // cofactor-transpose (adjugate)
    c(0, 0) = m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1);
    c(1, 0) = m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2);
    c(2, 0) = m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0);
    c(0, 1) = m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2);
    c(1, 1) = m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0);
    c(2, 1) = m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1);
    c(0, 2) = m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1);
    c(1, 2) = m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2);
    c(2, 2) = m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0);
    const double det = c(0, 0) * m(0, 0) + c(1, 0) * m(0, 1) + c(2, 0) * m(0, 2);
    const double invdet = 1.0 / det;
    return c * invdet;
}
inline double frob_norm(const Mat3 &m) {
    double s = 0;
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) s += m(i, j) * m(i, j);
    return std::sqrt(s);
}

// ---- quaternion.h:36-103 (q = w,x,y,z) -------------------------------------------------------
// Eigen::Quaterniond(w,x,y,z).toRotationMatrix()  (quaternion.h:36-38)
inline Mat3 quat_to_rotmat(const Vec4 &q) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    Mat3 R;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz;       R(0, 2) = txz + twy;
    R(1, 0) = txy + twz;       R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy;       R(2, 1) = tyz + twx;       R(2, 2) = 1 - (txx + tyy);
    return R;
}
// Eigen::Quaterniond(R), not yet normalised (w,x,y,z)  (quaternion.h:45-48)
inline Vec4 rotmat_to_quat_raw(const Mat3 &R) {
    double q[4]; // x,y,z at [0..2], w at [3]
    double t = R(0, 0) + R(1, 1) + R(2, 2);
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R(2, 1) - R(1, 2)) * t;
        q[1] = (R(0, 2) - R(2, 0)) * t;
        q[2] = (R(1, 0) - R(0, 1)) * t;
    } else {
        int i = 0;
        if (R(1, 1) > R(0, 0)) i = 1;
        if (R(2, 2) > R(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R(k, j) - R(j, k)) * t;
        q[j] = (R(j, i) + R(i, j)) * t;
        q[k] = (R(k, i) + R(i, k)) * t;
    }
    Vec4 out;
    out[0] = q[3]; out[1] = q[0]; out[2] = q[1]; out[3] = q[2];
    return out;
}
// Eigen::Quaterniond(R) then q.normalize()  (quaternion.h:45-51)
inline Vec4 rotmat_to_quat(const Mat3 &R) {
    Vec4 out = rotmat_to_quat_raw(R);
    double n2 = out[0] * out[0] + out[1] * out[1] + out[2] * out[2] + out[3] * out[3];
    if (n2 > 0) {
        double n = std::sqrt(n2);
        for (int a = 0; a < 4; ++a) out[a] /= n;
    }
    return out;
}
// quaternion.h:52-59
inline Vec4 quat_multiply(const Vec4 &qa, const Vec4 &qb) {
    const double a1 = qa[0], a2 = qa[1], a3 = qa[2], a4 = qa[3];
    const double b1 = qb[0], b2 = qb[1], b3 = qb[2], b4 = qb[3];
    Vec4 r;
    r[0] = a1 * b1 - a2 * b2 - a3 * b3 - a4 * b4;
    r[1] = a1 * b2 + a2 * b1 + a3 * b4 - a4 * b3;
    r[2] = a1 * b3 + a3 * b1 - a2 * b4 + a4 * b2;
    r[3] = a1 * b4 + a2 * b3 - a3 * b2 + a4 * b1;
    return r;
}
// quaternion.h:61-70
inline Vec3 quat_rotate(const Vec4 &q, const Vec3 &p) {
    const double q1 = q[0], q2 = q[1], q3 = q[2], q4 = q[3];
    const double p1 = p[0], p2 = p[1], p3 = p[2];
    const double px1 = -p1 * q2 - p2 * q3 - p3 * q4;
    const double px2 = p1 * q1 - p2 * q4 + p3 * q3;
    const double px3 = p2 * q1 + p1 * q4 - p3 * q2;
    const double px4 = p2 * q2 - p1 * q3 + p3 * q1;
    return mk3(px2 * q1 - px1 * q2 - px3 * q4 + px4 * q3, px3 * q1 - px1 * q3 + px2 * q4 - px4 * q2,
               px3 * q2 - px2 * q3 - px1 * q4 + px4 * q1);
}
// quaternion.h:73-96
inline Vec4 quat_exp(const Vec3 &w) {
    const double theta2 = sqnorm(w);
    const double theta = std::sqrt(theta2);
    const double theta_half = 0.5 * theta;
    double re, im;
    if (theta > 1e-6) {
        re = std::cos(theta_half);
        im = std::sin(theta_half) / theta;
    } else {
        const double theta4 = theta2 * theta2;
        re = 1.0 - (1.0 / 8.0) * theta2 + (1.0 / 384.0) * theta4;
        im = 0.5 - (1.0 / 48.0) * theta2 + (1.0 / 3840.0) * theta4;
        const double s = std::sqrt(re * re + im * im * theta2);
        re /= s;
        im /= s;
    }
    Vec4 r;
    r[0] = re; r[1] = im * w[0]; r[2] = im * w[1]; r[3] = im * w[2];
    return r;
}
inline Vec4 quat_step_pre(const Vec4 &q, const Vec3 &w) { return quat_multiply(quat_exp(w), q); }
inline Vec4 quat_step_post(const Vec4 &q, const Vec3 &w) { return quat_multiply(q, quat_exp(w)); }

// camera_pose.h:40-68
struct CameraPose {
    Vec4 q;
    Vec3 t;
    CameraPose() { q[0] = 1; q[1] = q[2] = q[3] = 0; t[0] = t[1] = t[2] = 0; }
    Mat3 R() const { return quat_to_rotmat(q); }
    Vec3 rotate(const Vec3 &p) const { return quat_rotate(q, p); }
};
inline CameraPose pose_from_Rt(const Mat3 &R, const Vec3 &t) {
    CameraPose p;
    p.q = rotmat_to_quat(R);
    p.t = t;
    return p;
}

// ---- Eigen dense decompositions restated ---------------------------------

// fullPivHouseholderQr().matrixQ() for a ROWS x COLS matrix A (column-major a[c*ROWS+r] on input,
// destroyed).  Writes Q (ROWS x ROWS, column-major q[c*ROWS+r]).
template <int ROWS, int COLS> inline void full_piv_householder_q(double *a, double *Q) {
    constexpr int size = (ROWS < COLS) ? ROWS : COLS;
    int rows_transp[size];
    double hcoeffs[size];
    const double precision = std::numeric_limits<double>::epsilon() * double(size);
    double biggest = 0.0;
    auto A = [&](int r, int c) -> double & { return a[c * ROWS + r]; };
    for (int k = 0; k < size; ++k) {
        // max |a_ij| over the trailing corner, first maximum in column-major order
        int rb = k, cb = k;
        double best = -1.0;
        for (int c = k; c < COLS; ++c)
            for (int r = k; r < ROWS; ++r) {
                double v = std::abs(A(r, c));
                if (v > best) { best = v; rb = r; cb = c; }
            }
        if (k == 0) biggest = best;
        if (std::abs(best) <= std::abs(biggest) * precision) {
            for (int i = k; i < size; ++i) { rows_transp[i] = i; hcoeffs[i] = 0.0; }
            break;
        }
        rows_transp[k] = rb;
        if (k != rb)
            for (int c = k; c < COLS; ++c) std::swap(A(k, c), A(rb, c));
        if (k != cb)
            for (int r = 0; r < ROWS; ++r) std::swap(A(r, k), A(r, cb));
        // makeHouseholderInPlace on col(k).tail(ROWS-k)
        double tail_sq = 0.0;
        for (int r = k + 1; r < ROWS; ++r) tail_sq += A(r, k) * A(r, k);
        const double c0 = A(k, k);
        double tau, beta;
        if (tail_sq <= std::numeric_limits<double>::min()) {
            tau = 0.0;
            beta = c0;
            for (int r = k + 1; r < ROWS; ++r) A(r, k) = 0.0;
        } else {
            beta = std::sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0) beta = -beta;
            for (int r = k + 1; r < ROWS; ++r) A(r, k) = A(r, k) / (c0 - beta);
            tau = (beta - c0) / beta;
        }
        hcoeffs[k] = tau;
        A(k, k) = beta;
        // apply H = I - tau v v^T (v = [1; essential]) to the trailing columns
        if (tau != 0.0) {
            for (int c = k + 1; c < COLS; ++c) {
                double tmp = 0.0;
                for (int r = k + 1; r < ROWS; ++r) tmp += A(r, k) * A(r, c);
                tmp += A(k, c);
                A(k, c) -= tau * tmp;
                for (int r = k + 1; r < ROWS; ++r) A(r, c) -= tau * A(r, k) * tmp;
            }
        }
    }
    // matrixQ: identity, then for k = size-1..0 apply H_k on block(k,k) and swap rows k <-> transp[k]
    for (int c = 0; c < ROWS; ++c)
        for (int r = 0; r < ROWS; ++r) Q[c * ROWS + r] = (r == c) ? 1.0 : 0.0;
    auto QQ = [&](int r, int c) -> double & { return Q[c * ROWS + r]; };
    for (int k = size - 1; k >= 0; --k) {
        const double tau = hcoeffs[k];
        if (tau != 0.0) {
            for (int c = k; c < ROWS; ++c) {
                double tmp = 0.0;
                for (int r = k + 1; r < ROWS; ++r) tmp += A(r, k) * QQ(r, c);
                tmp += QQ(k, c);
                QQ(k, c) -= tau * tmp;
                for (int r = k + 1; r < ROWS; ++r) QQ(r, c) -= tau * A(r, k) * tmp;
            }
        }
        const int rt = rows_transp[k];
        if (rt != k)
            for (int c = 0; c < ROWS; ++c) std::swap(QQ(k, c), QQ(rt, c));
    }
}

// householderQr().householderQ() for a ROWS x COLS matrix (ROWS >= COLS, column-major a[c*ROWS+r], destroyed): the same
// reflectors as above without any pivoting (Eigen's HouseholderQR, unblocked for these sizes).  Q is ROWS x ROWS,
// column-major.  relpose_8pt.cc:65 takes its last column as the nullspace of the 8 x 9 epipolar system.
template <int ROWS, int COLS> inline void householder_q(double *a, double *Q) {
    constexpr int size = (ROWS < COLS) ? ROWS : COLS;
    double hcoeffs[size];
    auto A = [&](int r, int c) -> double & { return a[c * ROWS + r]; };
    for (int k = 0; k < size; ++k) {
        double tail_sq = 0.0;
        for (int r = k + 1; r < ROWS; ++r) tail_sq += A(r, k) * A(r, k);
        const double c0 = A(k, k);
        double tau, beta;
        if (tail_sq <= std::numeric_limits<double>::min()) {
            tau = 0.0;
            beta = c0;
            for (int r = k + 1; r < ROWS; ++r) A(r, k) = 0.0;
        } else {
            beta = std::sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0) beta = -beta;
            for (int r = k + 1; r < ROWS; ++r) A(r, k) = A(r, k) / (c0 - beta);
            tau = (beta - c0) / beta;
        }
        hcoeffs[k] = tau;
        A(k, k) = beta;
        if (tau != 0.0) {
            for (int c = k + 1; c < COLS; ++c) {
                double tmp = 0.0;
                for (int r = k + 1; r < ROWS; ++r) tmp += A(r, k) * A(r, c);
                tmp += A(k, c);
                A(k, c) -= tau * tmp;
                for (int r = k + 1; r < ROWS; ++r) A(r, c) -= tau * A(r, k) * tmp;
            }
        }
    }
    for (int c = 0; c < ROWS; ++c)
        for (int r = 0; r < ROWS; ++r) Q[c * ROWS + r] = (r == c) ? 1.0 : 0.0;
    auto QQ = [&](int r, int c) -> double & { return Q[c * ROWS + r]; };
    for (int k = size - 1; k >= 0; --k) {
        const double tau = hcoeffs[k];
        if (tau == 0.0) continue;
        for (int c = k; c < ROWS; ++c) {
            double tmp = 0.0;
            for (int r = k + 1; r < ROWS; ++r) tmp += A(r, k) * QQ(r, c);
            tmp += QQ(k, c);
            QQ(k, c) -= tau * tmp;
            for (int r = k + 1; r < ROWS; ++r) QQ(r, c) -= tau * A(r, k) * tmp;
        }
    }
}

// SelfAdjointEigenSolver<Matrix<double,N,N>>: eigenvalues ascending, eigenvectors as columns (column-major V[c*N+r]).
// Eigen tridiagonalises and runs implicit QL; this restatement is a cyclic Jacobi iteration — same eigenpairs up to the
// sign of each vector and to roundoff (an iterative method cannot be restated bit for bit; tolerance parity only).
// Only the lower triangle of the row-major input is read.
template <int N> inline void sym_eigen_jacobi(const double *Ain, double *evals, double *V) {
    double A[N][N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) A[i][j] = (j <= i) ? Ain[i * N + j] : Ain[j * N + i];
    for (int c = 0; c < N; ++c)
        for (int r = 0; r < N; ++r) V[c * N + r] = (r == c) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < N; ++i) {
            diag += A[i][i] * A[i][i];
            for (int j = i + 1; j < N; ++j) off += A[i][j] * A[i][j];
        }
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = ((theta >= 0) ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < N; ++k) { // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - sn * akq;
                    A[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) { // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - sn * aqk;
                    A[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {
                    const double vkp = V[p * N + k], vkq = V[q * N + k];
                    V[p * N + k] = c * vkp - sn * vkq;
                    V[q * N + k] = sn * vkp + c * vkq;
                }
            }
    }
    int idx[N];
    for (int i = 0; i < N; ++i) idx[i] = i;
    std::sort(idx, idx + N, [&](int a, int b) { return A[a][a] < A[b][b]; });
    double Vs[N * N];
    for (int c = 0; c < N; ++c) {
        evals[c] = A[idx[c]][idx[c]];
        for (int r = 0; r < N; ++r) Vs[c * N + r] = V[idx[c] * N + r];
    }
    std::memcpy(V, Vs, sizeof(Vs));
}

// partialPivLu().solve(B): A is n x n row-major (a[r*lda+c]); B is n x nrhs row-major, overwritten by X.
inline void partial_piv_lu_solve(int n, double *a, int lda, double *b, int ldb, int nrhs) {
    int piv[32];
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = std::abs(a[k * lda + k]);
        for (int r = k + 1; r < n; ++r) {
            double v = std::abs(a[r * lda + k]);
            if (v > best) { best = v; p = r; }
        }
        piv[k] = p;
        if (best != 0.0) {
            if (p != k)
                for (int c = 0; c < n; ++c) std::swap(a[k * lda + c], a[p * lda + c]);
            const double pv = a[k * lda + k];
            for (int r = k + 1; r < n; ++r) a[r * lda + k] /= pv;
        }
        for (int r = k + 1; r < n; ++r) {
            const double l = a[r * lda + k];
            for (int c = k + 1; c < n; ++c) a[r * lda + c] -= l * a[k * lda + c];
        }
    }
    for (int k = 0; k < n; ++k)
        if (piv[k] != k)
            for (int c = 0; c < nrhs; ++c) std::swap(b[k * ldb + c], b[piv[k] * ldb + c]);
    for (int c = 0; c < nrhs; ++c) {
        for (int r = 1; r < n; ++r) {
            double s = b[r * ldb + c];
            for (int k = 0; k < r; ++k) s -= a[r * lda + k] * b[k * ldb + c];
            b[r * ldb + c] = s;
        }
        for (int r = n - 1; r >= 0; --r) {
            double s = b[r * ldb + c];
            for (int k = r + 1; k < n; ++k) s -= a[r * lda + k] * b[k * ldb + c];
            b[r * ldb + c] = s / a[r * lda + r];
        }
    }
}

// colPivHouseholderQr().solve(b) for a 3x2 system (relpose_5pt.cc:381 fallback): least squares.
inline void col_piv_qr_solve_3x2(const double B[3][2], const double b[3], double x[2]) {
    double A[3][2], rhs[3];
    for (int i = 0; i < 3; ++i) { A[i][0] = B[i][0]; A[i][1] = B[i][1]; rhs[i] = b[i]; }
    int perm[2] = {0, 1};
    double n0 = A[0][0] * A[0][0] + A[1][0] * A[1][0] + A[2][0] * A[2][0];
    double n1 = A[0][1] * A[0][1] + A[1][1] * A[1][1] + A[2][1] * A[2][1];
    if (n1 > n0) {
        for (int i = 0; i < 3; ++i) std::swap(A[i][0], A[i][1]);
        std::swap(perm[0], perm[1]);
    }
    for (int k = 0; k < 2; ++k) {
        double tail_sq = 0.0;
        for (int r = k + 1; r < 3; ++r) tail_sq += A[r][k] * A[r][k];
        const double c0 = A[k][k];
        double tau, beta;
        double ess[3] = {0, 0, 0};
        if (tail_sq <= std::numeric_limits<double>::min()) {
            tau = 0.0;
            beta = c0;
        } else {
            beta = std::sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0) beta = -beta;
            for (int r = k + 1; r < 3; ++r) ess[r] = A[r][k] / (c0 - beta);
            tau = (beta - c0) / beta;
        }
        A[k][k] = beta;
        for (int r = k + 1; r < 3; ++r) A[r][k] = 0.0;
        if (tau != 0.0) {
            for (int c = k + 1; c < 2; ++c) {
                double tmp = A[k][c];
                for (int r = k + 1; r < 3; ++r) tmp += ess[r] * A[r][c];
                A[k][c] -= tau * tmp;
                for (int r = k + 1; r < 3; ++r) A[r][c] -= tau * ess[r] * tmp;
            }
            double tmp = rhs[k];
            for (int r = k + 1; r < 3; ++r) tmp += ess[r] * rhs[r];
            rhs[k] -= tau * tmp;
            for (int r = k + 1; r < 3; ++r) rhs[r] -= tau * ess[r] * tmp;
        }
    }
    double y1 = rhs[1] / A[1][1];
    double y0 = (rhs[0] - A[0][1] * y1) / A[0][0];
    x[perm[0]] = y0;
    x[perm[1]] = y1;
}

// selfadjointView<Lower>().llt().solve(rhs): only the lower triangle of A (n x n row-major) is read.
// Returns false if a pivot is not positive (Eigen would carry NaNs on; we do too but flag it).
inline bool llt_solve_lower(int n, const double *A, const double *rhs, double *x) { // n <= 32
    double L[32 * 32];
    bool ok = true;
    for (int j = 0; j < n; ++j) {
        double d = A[j * n + j];
        for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0)) ok = false;
        const double ljj = std::sqrt(d);
        L[j * n + j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            double s = A[i * n + j];
            for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = s / ljj;
        }
    }
    double y[32];
    for (int i = 0; i < n; ++i) {
        double s = rhs[i];
        for (int k = 0; k < i; ++k) s -= L[i * n + k] * y[k];
        y[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * x[k];
        x[i] = s / L[i * n + i];
    }
    return ok;
}

// JacobiSVD<Matrix3d>(F, FullU|FullV): one-sided Jacobi on columns, singular values sorted descending.
// (optim_utils.h:59-73 only needs U, V and s up to joint column sign.)
inline void svd3(const Mat3 &F, Mat3 &U, double s[3], Mat3 &V) {
    Mat3 A = F;
    V = mat3_identity();
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) {
                    alpha += A(i, p) * A(i, p);
                    beta += A(i, q) * A(i, q);
                    gamma += A(i, p) * A(i, q);
                }
                if (gamma == 0.0) continue;
                off = std::max(off, std::abs(gamma) / std::sqrt(alpha * beta + 1e-300));
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = ((zeta >= 0) ? 1.0 : -1.0) / (std::abs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
                for (int i = 0; i < 3; ++i) {
                    const double ap = A(i, p), aq = A(i, q);
                    A(i, p) = c * ap - sn * aq;
                    A(i, q) = sn * ap + c * aq;
                    const double vp = V(i, p), vq = V(i, q);
                    V(i, p) = c * vp - sn * vq;
                    V(i, q) = sn * vp + c * vq;
                }
            }
        if (off < 1e-16) break;
    }
    double sv[3];
    for (int j = 0; j < 3; ++j) sv[j] = norm(col(A, j));
    int idx[3] = {0, 1, 2};
    std::sort(idx, idx + 3, [&](int a, int b) { return sv[a] > sv[b]; });
    Mat3 Vs, Us;
    for (int j = 0; j < 3; ++j) {
        s[j] = sv[idx[j]];
        set_col(Vs, j, col(V, idx[j]));
        if (s[j] > 0) set_col(Us, j, col(A, idx[j]) / s[j]);
    }
    // complete U for (near-)zero singular values (rank-2 F): u2 = u0 x u1
    if (!(s[2] > 1e-14 * s[0])) set_col(Us, 2, cross(col(Us, 0), col(Us, 1)));
    if (!(s[1] > 0)) { // degenerate rank-1 input; any orthonormal completion
        Vec3 u0 = col(Us, 0);
        Vec3 e = (std::abs(u0[0]) < 0.9) ? mk3(1, 0, 0) : mk3(0, 1, 0);
        Vec3 u1 = normalized(cross(u0, e));
        set_col(Us, 1, u1);
        set_col(Us, 2, cross(u0, u1));
    }
    U = Us;
    V = Vs;
}

} // namespace plo
