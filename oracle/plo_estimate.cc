// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// The four robust.cc entry points for PINHOLE-family cameras (paths relative to /root/reference).
#include "plo.h"

namespace plo {

namespace {
// camera_models.cc:668-711 (PINHOLE unproject): ((x-cx)/fx, (y-cy)/fy, 1) then hnormalized (camera_models.h:98-102)
inline Vec2 unproject(const SimpleCamera &cam, const Vec2 &p) {
    Vec2 r;
    r[0] = (p[0] - cam.cx) / cam.fx;
    r[1] = (p[1] - cam.cy) / cam.fy;
    return r;
}
} // namespace

// robust.cc:36-126 (branch without focal estimation)
RansacStats estimate_absolute_pose(const std::vector<Vec2> &points2D, const std::vector<Vec3> &points3D,
                                   const RansacOptions &ropt, const BundleOptions &bopt, double max_error,
                                   const SimpleCamera &cam, CameraPose *pose, std::vector<char> *inliers,
                                   Counters *cnt) {
    std::vector<Vec2> norm(points2D.size());
    for (size_t k = 0; k < points2D.size(); ++k) norm[k] = unproject(cam, points2D[k]);
    double scale = 1.0 / cam.focal();
    const double max_error_scaled = max_error * scale;
    RansacStats stats = ransac_pnp(norm, points3D, ropt, max_error_scaled, pose, inliers, cnt);
    if (stats.num_inliers > 3) { // :103-123
        std::vector<Vec2> x_in;
        std::vector<Vec3> X_in;
        scale = 1.0 / cam.focal();
        BundleOptions b = bopt;
        b.loss_scale = bopt.loss_scale * scale;
        for (size_t k = 0; k < points2D.size(); ++k) {
            if (!(*inliers)[k]) continue;
            Vec2 p;
            p[0] = points2D[k][0] * scale;
            p[1] = points2D[k][1] * scale;
            x_in.push_back(p);
            X_in.push_back(points3D[k]);
        }
        // camera.rescale(scale): focal and principal point multiplied by scale (camera_models.cc:432-454)
        SimpleCamera rc = cam;
        rc.fx *= scale; rc.fy *= scale; rc.cx *= scale; rc.cy *= scale;
        bundle_adjust_camera(x_in, X_in, rc, pose, b);
    }
    return stats;
}

// robust.cc:242-314 (tangent_sampson == false branch)
RansacStats estimate_relative_pose(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const SimpleCamera &cam1,
                                   const SimpleCamera &cam2, const RansacOptions &ropt, const BundleOptions &bopt,
                                   double max_error, CameraPose *pose, std::vector<char> *inliers, Counters *cnt) {
    const size_t n = x1.size();
    const double scale = 0.5 * (1.0 / cam1.focal() + 1.0 / cam2.focal());
    const double max_error_scaled = max_error * scale;
    BundleOptions b = bopt;
    b.loss_scale *= scale;
    std::vector<Vec2> c1(n), c2(n);
    for (size_t k = 0; k < n; ++k) {
        c1[k] = unproject(cam1, x1[k]);
        c2[k] = unproject(cam2, x2[k]);
    }
    RansacStats stats = ransac_relpose(c1, c2, ropt, max_error_scaled, pose, inliers, cnt);
    if (stats.num_inliers > 5) {
        std::vector<Vec2> a, bb;
        a.reserve(stats.num_inliers);
        bb.reserve(stats.num_inliers);
        for (size_t k = 0; k < n; ++k) {
            if (!(*inliers)[k]) continue;
            a.push_back(c1[k]);
            bb.push_back(c2[k]);
        }
        refine_relpose(a, bb, pose, b);
    }
    return stats;
}

// robust.cc:544-594
RansacStats estimate_fundamental(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const RansacOptions &ropt,
                                 const BundleOptions &bopt, double max_error, bool real_focal_check, Mat3 *F,
                                 std::vector<char> *inliers, Counters *cnt) {
    const size_t n = x1.size();
    if (n < 7) return RansacStats();
    Mat3 T1, T2;
    std::vector<Vec2> a = x1, b = x2;
    const double scale = normalize_points(a, b, T1, T2, true, !real_focal_check, true);
    const double max_error_scaled = max_error / scale;
    BundleOptions bo = bopt;
    bo.loss_scale /= scale;
    if (ropt.score_initial_model) { // :566-569
        *F = inverse3(transpose(T2)) * (*F) * inverse3(T1);
        *F = *F * (1.0 / frob_norm(*F));
    }
    RansacStats stats = ransac_fundamental(a, b, ropt, max_error_scaled, real_focal_check, F, inliers, cnt);
    if (stats.num_inliers > 7) {
        std::vector<Vec2> ia, ib;
        for (size_t k = 0; k < n; ++k) {
            if (!(*inliers)[k]) continue;
            ia.push_back(a[k]);
            ib.push_back(b[k]);
        }
        refine_fundamental(ia, ib, F, bo);
    }
    *F = transpose(T2) * (*F) * T1;
    const double nf = frob_norm(*F);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) (*F)(r, c) /= nf;
    return stats;
}

// robust.cc:712-757
RansacStats estimate_homography(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const RansacOptions &ropt,
                                const BundleOptions &bopt, double max_error, Mat3 *H, std::vector<char> *inliers,
                                Counters *cnt) {
    const size_t n = x1.size();
    if (n < 4) return RansacStats();
    Mat3 T1, T2;
    std::vector<Vec2> a = x1, b = x2;
    const double scale = normalize_points(a, b, T1, T2, true, true, true);
    const double max_error_scaled = max_error / scale;
    BundleOptions bo = bopt;
    bo.loss_scale /= scale;
    if (ropt.score_initial_model) { // :729-732
        *H = T2 * (*H) * inverse3(T1);
        *H = *H * (1.0 / frob_norm(*H));
    }
    RansacStats stats = ransac_homography(a, b, ropt, max_error_scaled, H, inliers, cnt);
    if (stats.num_inliers > 4) {
        std::vector<Vec2> ia, ib;
        for (size_t k = 0; k < n; ++k) {
            if (!(*inliers)[k]) continue;
            ia.push_back(a[k]);
            ib.push_back(b[k]);
        }
        refine_homography(ia, ib, H, bo);
    }
    *H = inverse3(T2) * (*H) * T1;
    const double nh = frob_norm(*H);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) (*H)(r, c) /= nh;
    return stats;
}

} // namespace plo
