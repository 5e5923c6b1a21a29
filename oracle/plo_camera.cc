// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY PARTLY PINNED: sampler, loop control flow, iteration arithmetic, univariate / p3p scalar solvers, Sturm root isolation, F / H scorers, masks, the real-focal check and the scalar camera code against the reference's own code (oracle/_ref, oracle/ref/ref_capi.cc); the transcription of PoseLib's logic for the WHOLE path (solvers, scorers, refiners, estimators, estimate_*) against the reference's own sources run on mini-Eigen (oracle/_ref/libplref2.so, oracle/ref/ref2_capi.cc, tests/test_ref_sources.py); Eigen's own arithmetic (reduction order, decompositions) is UNPINNED.
// Camera models on the path of estimate_absolute_pose / estimate_relative_pose
// (misc/camera_models.{h,cc}, paths relative to /root/reference/PoseLib):
// NULL, SIMPLE_PINHOLE, PINHOLE, SIMPLE_RADIAL, RADIAL, OPENCV.
#include "plo.h"
#include <stdexcept>

namespace plo {

namespace {
const double UNDIST_TOL = 1e-10;  // camera_models.cc:41
const size_t UNDIST_MAX_ITER = 100; // camera_models.cc:42

// camera_models.cc:579-595
double undistort_poly1(double k1, double rd) {
    double r = rd;
    for (size_t iter = 0; iter < UNDIST_MAX_ITER; ++iter) {
        double r2 = r * r;
        double f = k1 * r2 * r + r - rd;
        if (std::abs(f) < UNDIST_TOL) break;
        double fp = 3.0 * k1 * r2 + 1.0;
        r = r - f / fp;
    }
    return r;
}
// camera_models.cc:597-613
double undistort_poly2(double k1, double k2, double rd) {
    double r = rd;
    for (size_t iter = 0; iter < UNDIST_MAX_ITER; ++iter) {
        double r2 = r * r;
        double f = k2 * r2 * r2 * r + k1 * r2 * r + r - rd;
        if (std::abs(f) < UNDIST_TOL) break;
        double fp = 5.0 * k2 * r2 * r2 + 3.0 * k1 * r2 + 1.0;
        r = r - f / fp;
    }
    return r;
}
// camera_models.cc:919-931
void opencv_distortion(double k1, double k2, double p1, double p2, const double x[2], double xp[2]) {
    const double u = x[0], v = x[1];
    const double u2 = u * u, uv = u * v, v2 = v * v;
    const double r2 = u * u + v * v;
    const double alpha = 1.0 + k1 * r2 + k2 * r2 * r2;
    xp[0] = alpha * u + 2.0 * p1 * uv + p2 * (r2 + 2.0 * u2);
    xp[1] = alpha * v + 2.0 * p2 * uv + p1 * (r2 + 2.0 * v2);
}
// camera_models.cc:933-950 (point Jacobian only)
void opencv_distortion_jac(double k1, double k2, double p1, double p2, const double x[2], double xp[2],
                           double jac[2][2]) {
    const double u = x[0], v = x[1];
    const double u2 = u * u, uv = u * v, v2 = v * v;
    const double r2 = u * u + v * v;
    jac[0][0] = k2 * r2 * r2 + 6 * p2 * u + 2 * p1 * v + u * (2 * k1 * u + 4 * k2 * u * r2) + k1 * r2 + 1.0;
    jac[0][1] = 2 * p1 * u + 2 * p2 * v + v * (2 * k1 * u + 4 * k2 * u * r2);
    jac[1][0] = 2 * p1 * u + 2 * p2 * v + u * (2 * k1 * v + 4 * k2 * v * r2);
    jac[1][1] = k2 * r2 * r2 + 2 * p2 * u + 6 * p1 * v + v * (2 * k1 * v + 4 * k2 * v * r2) + k1 * r2 + 1.0;
    const double alpha = 1.0 + k1 * r2 + k2 * r2 * r2;
    xp[0] = alpha * u + 2.0 * p1 * uv + p2 * (r2 + 2.0 * u2);
    xp[1] = alpha * v + 2.0 * p2 * uv + p1 * (r2 + 2.0 * v2);
}
// camera_models.cc:972-991; the 2x2 inverse is Eigen's closed form (Inverse_impl size 2: 1/det times the adjugate)
void undistort_opencv(double k1, double k2, double p1, double p2, const double xp[2], double x[2]) {
    x[0] = xp[0];
    x[1] = xp[1];
    const double lambda = 1e-8;
    for (size_t iter = 0; iter < UNDIST_MAX_ITER; ++iter) {
        double xd[2], jac[2][2];
        opencv_distortion_jac(k1, k2, p1, p2, x, xd, jac);
        jac[0][0] += lambda;
        jac[1][1] += lambda;
        const double res0 = xd[0] - xp[0], res1 = xd[1] - xp[1];
        if (std::sqrt(res0 * res0 + res1 * res1) < UNDIST_TOL) break;
        const double det = jac[0][0] * jac[1][1] - jac[1][0] * jac[0][1];
        const double invdet = 1.0 / det;
        const double i00 = jac[1][1] * invdet, i01 = -jac[0][1] * invdet;
        const double i10 = -jac[1][0] * invdet, i11 = jac[0][0] * invdet;
        x[0] = x[0] - (i00 * res0 + i01 * res1);
        x[1] = x[1] - (i10 * res0 + i11 * res1);
    }
}
inline void normalize3(Vec3 *x) { // Eigen normalize(): divide by sqrt(squaredNorm) when > 0
    const double n2 = (*x)[0] * (*x)[0] + (*x)[1] * (*x)[1] + (*x)[2] * (*x)[2];
    if (n2 > 0) {
        const double n = std::sqrt(n2);
        (*x)[0] /= n;
        (*x)[1] /= n;
        (*x)[2] /= n;
    }
}
// jac = jac_d (2x2) * jac (2x3)
inline void left_mul_2x2(const double jd[2][2], double jac[2][3]) {
    double o[2][3];
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) o[r][c] = jd[r][0] * jac[0][c] + jd[r][1] * jac[1][c];
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) jac[r][c] = o[r][c];
}
} // namespace

int Camera::num_params() const {
    switch (model_id) {
    case CAM_NULL: return 0;
    case CAM_SIMPLE_PINHOLE: return 3;
    case CAM_PINHOLE: return 4;
    case CAM_SIMPLE_RADIAL: return 4;
    case CAM_RADIAL: return 5;
    case CAM_OPENCV: return 8;
    default: throw std::runtime_error("NYI");
    }
}

// camera_models.cc:304-324: mean of the focal_idx entries (accumulated from 0.0), 1.0 for the empty camera
double Camera::focal() const {
    switch (model_id) {
    case CAM_NULL: return 1.0;
    case CAM_SIMPLE_PINHOLE:
    case CAM_SIMPLE_RADIAL:
    case CAM_RADIAL: return 0.0 + params[0] / 1;
    case CAM_PINHOLE:
    case CAM_OPENCV: return 0.0 + params[0] / 2 + params[1] / 2;
    default: throw std::runtime_error("NYI");
    }
}

// camera_models.cc:432-454: focal_idx then principal_point_idx entries multiplied by scale
void Camera::rescale(double scale) {
    switch (model_id) {
    case CAM_NULL: return;
    case CAM_SIMPLE_PINHOLE:
    case CAM_SIMPLE_RADIAL:
    case CAM_RADIAL:
        params[0] *= scale;
        params[1] *= scale;
        params[2] *= scale;
        return;
    case CAM_PINHOLE:
    case CAM_OPENCV:
        params[0] *= scale;
        params[1] *= scale;
        params[2] *= scale;
        params[3] *= scale;
        return;
    default: throw std::runtime_error("NYI");
    }
}

// <Model>::project (camera_models.cc:668-671,716-720,766-772,849-857,965-970,2708-2710)
void Camera::project(const Vec3 &x, Vec2 *xp) const {
    const double *p = params;
    switch (model_id) {
    case CAM_NULL:
        (*xp)[0] = x[0] / x[2];
        (*xp)[1] = x[1] / x[2];
        return;
    case CAM_SIMPLE_PINHOLE:
        (*xp)[0] = p[0] * x[0] / x[2] + p[1];
        (*xp)[1] = p[0] * x[1] / x[2] + p[2];
        return;
    case CAM_PINHOLE:
        (*xp)[0] = p[0] * x[0] / x[2] + p[2];
        (*xp)[1] = p[1] * x[1] / x[2] + p[3];
        return;
    case CAM_SIMPLE_RADIAL: {
        const double inv_z = 1.0 / x[2];
        const double px = x[0] * inv_z, py = x[1] * inv_z;
        const double r2 = px * px + py * py;
        const double alpha = (1.0 + p[3] * r2);
        (*xp)[0] = p[0] * alpha * px + p[1];
        (*xp)[1] = p[0] * alpha * py + p[2];
        return;
    }
    case CAM_RADIAL: {
        const double hx = x[0] / x[2], hy = x[1] / x[2]; // hnormalized
        const double r2 = hx * hx + hy * hy;
        const double alpha = (1.0 + p[3] * r2 + p[4] * r2 * r2);
        (*xp)[0] = p[0] * alpha * hx + p[1];
        (*xp)[1] = p[0] * alpha * hy + p[2];
        return;
    }
    case CAM_OPENCV: {
        const double x0[2] = {x[0] / x[2], x[1] / x[2]};
        double d[2];
        opencv_distortion(p[4], p[5], p[6], p[7], x0, d);
        (*xp)[0] = p[0] * d[0] + p[2];
        (*xp)[1] = p[1] * d[1] + p[3];
        return;
    }
    default: throw std::runtime_error("NYI");
    }
}

// <Model>::project_with_jac, point Jacobian only
// (camera_models.cc:672-699,721-744,773-818,858-893,992-1024,2711-2723)
void Camera::project_with_jac(const Vec3 &x, Vec2 *xp, double jac[2][3]) const {
    const double *p = params;
    switch (model_id) {
    case CAM_NULL: {
        (*xp)[0] = x[0] / x[2];
        (*xp)[1] = x[1] / x[2];
        const double z_inv = 1.0 / x[2];
        jac[0][0] = z_inv; jac[0][1] = 0.0; jac[0][2] = -(*xp)[0] * z_inv;
        jac[1][0] = 0.0; jac[1][1] = z_inv; jac[1][2] = -(*xp)[1] * z_inv;
        return;
    }
    case CAM_SIMPLE_PINHOLE:
    case CAM_PINHOLE: {
        const double fx = p[0], fy = model_id == CAM_PINHOLE ? p[1] : p[0];
        const double cx = model_id == CAM_PINHOLE ? p[2] : p[1], cy = model_id == CAM_PINHOLE ? p[3] : p[2];
        const double inv_z = 1.0 / x[2];
        const double px = fx * x[0] * inv_z, py = fy * x[1] * inv_z;
        (*xp)[0] = px + cx;
        (*xp)[1] = py + cy;
        jac[0][0] = fx * inv_z; jac[0][1] = 0.0; jac[0][2] = -px * inv_z;
        jac[1][0] = 0.0; jac[1][1] = fy * inv_z; jac[1][2] = -py * inv_z;
        return;
    }
    case CAM_SIMPLE_RADIAL:
    case CAM_RADIAL: {
        const double inv_z = 1.0 / x[2];
        const double px = x[0] * inv_z, py = x[1] * inv_z;
        const double r2 = px * px + py * py;
        double jd[2][2], alpha;
        if (model_id == CAM_SIMPLE_RADIAL) {
            alpha = (1.0 + p[3] * r2);
            jd[0][0] = (2.0 * p[3] * px * px + alpha) * p[0];
            jd[0][1] = (2.0 * p[3] * px * py) * p[0];
            jd[1][0] = jd[0][1];
            jd[1][1] = (2.0 * p[3] * py * py + alpha) * p[0];
        } else {
            alpha = (1.0 + p[3] * r2 + p[4] * r2 * r2);
            const double alphap = (2.0 * p[3] + 4.0 * p[4] * r2);
            jd[0][0] = (alphap * px * px + alpha) * p[0];
            jd[0][1] = (alphap * px * py) * p[0];
            jd[1][0] = jd[0][1];
            jd[1][1] = (alphap * py * py + alpha) * p[0];
        }
        jac[0][0] = inv_z; jac[0][1] = 0; jac[0][2] = -px * inv_z;
        jac[1][0] = 0; jac[1][1] = inv_z; jac[1][2] = -py * inv_z;
        left_mul_2x2(jd, jac);
        (*xp)[0] = p[0] * alpha * px + p[1];
        (*xp)[1] = p[0] * alpha * py + p[2];
        return;
    }
    case CAM_OPENCV: {
        const double x0[2] = {x[0] / x[2], x[1] / x[2]};
        double d[2], jac0[2][2];
        opencv_distortion_jac(p[4], p[5], p[6], p[7], x0, d, jac0);
        jac[0][0] = 1.0 / x[2]; jac[0][1] = 0.0; jac[0][2] = -x0[0] / x[2];
        jac[1][0] = 0.0; jac[1][1] = 1.0 / x[2]; jac[1][2] = -x0[1] / x[2];
        left_mul_2x2(jac0, jac);
        for (int c = 0; c < 3; ++c) {
            jac[0][c] *= p[0];
            jac[1][c] *= p[1];
        }
        (*xp)[0] = p[0] * d[0] + p[2];
        (*xp)[1] = p[1] * d[1] + p[3];
        return;
    }
    default: throw std::runtime_error("NYI");
    }
}

// <Model>::unproject (camera_models.cc:700-705,745-751,819-831,895-907,1025-1033,2724-2726)
void Camera::unproject(const Vec2 &xp, Vec3 *x) const {
    const double *p = params;
    switch (model_id) {
    case CAM_NULL:
        *x = mk3(xp[0], xp[1], 1.0);
        return;
    case CAM_SIMPLE_PINHOLE:
        *x = mk3((xp[0] - p[1]) / p[0], (xp[1] - p[2]) / p[0], 1.0);
        normalize3(x);
        return;
    case CAM_PINHOLE:
        *x = mk3((xp[0] - p[2]) / p[0], (xp[1] - p[3]) / p[1], 1.0);
        normalize3(x);
        return;
    case CAM_SIMPLE_RADIAL:
    case CAM_RADIAL: {
        *x = mk3((xp[0] - p[1]) / p[0], (xp[1] - p[2]) / p[0], 0.0);
        const double r0 = std::sqrt((*x)[0] * (*x)[0] + (*x)[1] * (*x)[1] + (*x)[2] * (*x)[2]);
        if (std::abs(r0) > 1e-8) {
            const double r = model_id == CAM_SIMPLE_RADIAL ? undistort_poly1(p[3], r0) : undistort_poly2(p[3], p[4], r0);
            const double s = r / r0;
            (*x)[0] *= s;
            (*x)[1] *= s;
            (*x)[2] *= s;
        }
        (*x)[2] = 1.0;
        normalize3(x);
        return;
    }
    case CAM_OPENCV: {
        const double xp0[2] = {(xp[0] - p[2]) / p[0], (xp[1] - p[3]) / p[1]};
        double x0[2];
        undistort_opencv(p[4], p[5], p[6], p[7], xp0, x0);
        *x = mk3(x0[0], x0[1], 1.0);
        normalize3(x);
        return;
    }
    default: throw std::runtime_error("NYI");
    }
}

// camera_models.h:98-102: 3D unproject followed by hnormalized()
Vec2 Camera::unproject2(const Vec2 &xp) const {
    Vec3 x3;
    unproject(xp, &x3);
    Vec2 r;
    r[0] = x3[0] / x3[2];
    r[1] = x3[1] / x3[2];
    return r;
}

// camera_models.cc:456-489 (the default unproject_with_jac shared by all six models): M = J^T (J J^T)^-1
void Camera::unproject_with_jac(const Vec2 &xp, Vec3 *x, double M[3][2]) const {
    unproject(xp, x);
    double J[2][3];
    Vec2 xp_proj;
    project_with_jac(*x, &xp_proj, J);
    double B[2][2];
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) B[r][c] = J[r][0] * J[c][0] + J[r][1] * J[c][1] + J[r][2] * J[c][2];
    const double det = B[0][0] * B[1][1] - B[0][1] * B[1][0];
    double Bi[2][2] = {{B[1][1], -B[0][1]}, {-B[1][0], B[0][0]}};
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) Bi[r][c] /= det;
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 2; ++c) M[i][c] = J[0][i] * Bi[0][c] + J[1][i] * Bi[1][c];
}

} // namespace plo

// test hooks for the scalar helpers above (anonymous namespace), compared with the reference's camera_models.cc
// through oracle/_ref (tests/test_ref_pins.py)
extern "C" {
double plo_undistort_poly(double k1, double k2, int two, double rd) {
    return two ? plo::undistort_poly2(k1, k2, rd) : plo::undistort_poly1(k1, rd);
}
void plo_opencv_distortion(const double *d4, const double *x2, double *out2, double *jac4) {
    if (jac4) {
        double J[2][2];
        plo::opencv_distortion_jac(d4[0], d4[1], d4[2], d4[3], x2, out2, J);
        jac4[0] = J[0][0]; jac4[1] = J[0][1]; jac4[2] = J[1][0]; jac4[3] = J[1][1];
    } else {
        plo::opencv_distortion(d4[0], d4[1], d4[2], d4[3], x2, out2);
    }
}
void plo_camera_rescale(int model_id, double *params8, double scale) {
    plo::Camera c;
    c.model_id = model_id;
    for (int i = 0; i < 8; ++i) c.params[i] = params8[i];
    c.rescale(scale);
    for (int i = 0; i < 8; ++i) params8[i] = c.params[i];
}
}
