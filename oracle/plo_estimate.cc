// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY PARTLY PINNED: sampler, loop control flow, iteration arithmetic, univariate / p3p scalar solvers, Sturm root isolation, F / H scorers, masks, the real-focal check and the scalar camera code against the reference's own code (oracle/_ref, oracle/ref/ref_capi.cc); the transcription of PoseLib's logic for the WHOLE path (solvers, scorers, refiners, estimators, estimate_*) against the reference's own sources run on mini-Eigen (oracle/_ref/libplref2.so, oracle/ref/ref2_capi.cc, tests/test_ref_sources.py); Eigen's own arithmetic (reduction order, decompositions) is UNPINNED.
// The four robust.cc entry points (paths relative to /root/reference/PoseLib).
#include "plo.h"

namespace plo {

// robust.cc:36-126 (branch without focal estimation)
RansacStats estimate_absolute_pose(const std::vector<Vec2> &points2D, const std::vector<Vec3> &points3D,
                                   const RansacOptions &ropt, const BundleOptions &bopt, double max_error,
                                   const Camera &cam, CameraPose *pose, std::vector<char> *inliers, Counters *cnt) {
    std::vector<Vec2> norm(points2D.size());
    for (size_t k = 0; k < points2D.size(); ++k) norm[k] = cam.unproject2(points2D[k]);
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
        Camera rc = cam;
        rc.rescale(scale);
        bundle_adjust_camera(x_in, X_in, rc, pose, b);
    }
    return stats;
}

// robust.cc:242-314
RansacStats estimate_relative_pose(const std::vector<Vec2> &x1, const std::vector<Vec2> &x2, const Camera &cam1,
                                   const Camera &cam2, const RansacOptions &ropt, const BundleOptions &bopt,
                                   double max_error, CameraPose *pose, std::vector<char> *inliers, Counters *cnt,
                                   bool tangent_sampson) {
    const size_t n = x1.size();
    const double scale = 0.5 * (1.0 / cam1.focal() + 1.0 / cam2.focal());
    const double max_error_scaled = max_error * scale;
    BundleOptions b = bopt;
    b.loss_scale *= scale;
    if (tangent_sampson) { // :255-284
        std::vector<Vec2> s1(n), s2(n);
        for (size_t k = 0; k < n; ++k) {
            s1[k][0] = x1[k][0] * scale; s1[k][1] = x1[k][1] * scale;
            s2[k][0] = x2[k][0] * scale; s2[k][1] = x2[k][1] * scale;
        }
        Camera c1 = cam1, c2 = cam2;
        c1.rescale(scale);
        c2.rescale(scale);
        RansacStats stats = ransac_relpose(s1, s2, c1, c2, ropt, max_error_scaled, pose, inliers, cnt);
        if (stats.num_inliers > 5) {
            // refine_relpose(x1_inliers, x2_inliers, &pair, bundle) with fixed cameras (bundle.cc:237-247):
            // unproject_with_jac of the inlier points, then the FixCameraRelativePoseRefiner
            std::vector<Vec3> d1, d2;
            std::vector<Mat32> M1, M2;
            for (size_t k = 0; k < n; ++k) {
                if (!(*inliers)[k]) continue;
                Vec3 a, bb;
                Mat32 ma, mb;
                c1.unproject_with_jac(s1[k], &a, ma.m);
                c2.unproject_with_jac(s2[k], &bb, mb.m);
                d1.push_back(a); d2.push_back(bb);
                M1.push_back(ma); M2.push_back(mb);
            }
            refine_relpose(d1, d2, M1, M2, pose, b);
        }
        return stats;
    }
    std::vector<Vec2> c1(n), c2(n);
    for (size_t k = 0; k < n; ++k) {
        c1[k] = cam1.unproject2(x1[k]);
        c2[k] = cam2.unproject2(x2[k]);
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
