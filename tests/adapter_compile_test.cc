// Compile-and-link test of the PoseLib-signature adapter against minimal stand-ins for the Eigen / PoseLib types
// (Eigen is not installed in this image).  Built and run by tests/test_adapter.py; the run part needs a GPU.
#include "../poselib_b200/adapter/poselib_b200.hpp"

#include <array>
#include <cstdio>
#include <cstring>

namespace standin {
template <int N> struct Vec {
    double d[N];
    double &operator()(int i) { return d[i]; }
    double operator()(int i) const { return d[i]; }
};
using Vector2d = Vec<2>;
using Vector3d = Vec<3>;
using Vector4d = Vec<4>;
struct Matrix3d { // column-major
    double d[9];
};
struct CameraPose { // camera_pose.h:40-68
    Vector4d q{{1, 0, 0, 0}};
    Vector3d t{{0, 0, 0}};
};
struct Camera { // misc/camera_models.h:59-63
    int model_id = -1, width = 0, height = 0;
    std::vector<double> params;
};
struct Image {
    CameraPose pose;
    Camera camera;
};
struct RansacOptions { // types.h:39-50
    size_t max_iterations = 100000, min_iterations = 1000;
    double dyn_num_trials_mult = 3.0, success_prob = 0.9999;
    unsigned long seed = 0;
    bool progressive_sampling = false;
    size_t max_prosac_iterations = 100000;
    bool score_initial_model = false;
};
struct RansacStats { // types.h:52-58
    size_t refinements = 0, iterations = 0, num_inliers = 0;
    double inlier_ratio = 0, model_score = 0;
};
struct BundleOptions { // types.h:60-95
    size_t max_iterations = 100;
    enum LossType { TRIVIAL, TRUNCATED, HUBER, CAUCHY, TRUNCATED_CAUCHY, TRUNCATED_LE_ZACH } loss_type = CAUCHY;
    double loss_scale = 1.0, gradient_tol = 1e-12, step_tol = 1e-8, relative_cost_tol = 1e-10, initial_lambda = 1e-3,
           min_lambda = 1e-10, max_lambda = 1e10;
    bool verbose = false;
    enum LambdaUpdateType { NIELSEN, FIXED_FACTOR } lambda_update = NIELSEN;
    double lambda_factor = 10.0;
    enum DampingType { LEVENBERG, MARQUARDT } damping = LEVENBERG;
    bool refine_focal_length = false, refine_extra_params = false, refine_principal_point = false;
};
struct AbsolutePoseOptions {
    RansacOptions ransac;
    BundleOptions bundle;
    double max_error = 12.0;
    bool estimate_focal_length = false, estimate_extra_params = false;
};
struct RelativePoseOptions {
    RansacOptions ransac;
    BundleOptions bundle;
    double max_error = 1.0;
    bool tangent_sampson = false, real_focal_check = false;
};
struct HomographyOptions {
    RansacOptions ransac;
    BundleOptions bundle;
    double max_error = 1.0;
};
} // namespace standin

int main(int argc, char **argv) {
    using namespace standin;
    // instantiate every adapter entry point (link check)
    std::vector<Vector2d> x1(8), x2(8);
    std::vector<Vector3d> X(8);
    for (int i = 0; i < 8; ++i) {
        x1[i] = {{0.01 * i, 0.02 * i - 0.05}};
        x2[i] = {{0.011 * i + 0.01, 0.019 * i - 0.04}};
        X[i] = {{0.3 * i - 1.0, 0.2 * i, 4.0 + 0.1 * i}};
    }
    std::vector<char> inl;
    if (argc > 1 && std::strcmp(argv[1], "run") == 0) { // needs a GPU
        RelativePoseOptions ro;
        ro.ransac.max_iterations = 50;
        ro.ransac.min_iterations = 10;
        CameraPose pose;
        Camera cam;
        cam.model_id = 1;
        cam.params = {1.0, 1.0, 0.0, 0.0};
        RansacStats st = poselib_b200::estimate_relative_pose<RansacStats>(x1, x2, cam, cam, ro, &pose, &inl);
        std::printf("relpose iterations=%zu inliers=%zu\n", st.iterations, st.num_inliers);
        Matrix3d F{};
        st = poselib_b200::estimate_fundamental<RansacStats>(x1, x2, ro, &F, &inl);
        HomographyOptions ho;
        ho.ransac.max_iterations = 50;
        ho.ransac.min_iterations = 10;
        Matrix3d H{};
        st = poselib_b200::estimate_homography<RansacStats>(x1, x2, ho, &H, &inl);
        AbsolutePoseOptions ao;
        ao.ransac.max_iterations = 50;
        ao.ransac.min_iterations = 10;
        Image img;
        img.camera = cam;
        st = poselib_b200::estimate_absolute_pose<RansacStats>(x1, X, ao, &img, &inl);
        st = poselib_b200::ransac_pnp<RansacStats>(x1, X, ao, &pose, &inl);
        st = poselib_b200::ransac_relpose<RansacStats>(x1, x2, ro, &pose, &inl);
        Camera rad; // SIMPLE_RADIAL through the tangent-Sampson path
        rad.model_id = 2;
        rad.params = {1.0, 0.0, 0.0, -0.05};
        st = poselib_b200::ransac_relpose<RansacStats>(x1, x2, rad, rad, ro, &pose, &inl);
        ro.tangent_sampson = true;
        st = poselib_b200::estimate_relative_pose<RansacStats>(x1, x2, rad, rad, ro, &pose, &inl);
        ro.tangent_sampson = false;
        st = poselib_b200::ransac_fundamental<RansacStats>(x1, x2, ro, &F, &inl);
        st = poselib_b200::ransac_homography<RansacStats>(x1, x2, ho, &H, &inl);
        std::vector<Vector3d> b1(7), b2(7);
        for (int i = 0; i < 7; ++i) {
            b1[i] = {{x1[i](0), x1[i](1), 1.0}};
            b2[i] = {{x2[i](0), x2[i](1), 1.0}};
        }
        std::vector<CameraPose> poses;
        std::vector<Matrix3d> Ms;
        std::vector<Vector3d> b13(b1.begin(), b1.begin() + 3), X3(X.begin(), X.begin() + 3);
        poselib_b200::p3p(b13, X3, &poses);
        poselib_b200::p3p_lambdatwist(b13, X3, &poses);
        std::vector<Vector3d> b15(b1.begin(), b1.begin() + 5), b25(b2.begin(), b2.begin() + 5);
        poselib_b200::relpose_5pt(b15, b25, &Ms);
        poselib_b200::relpose_5pt_poses(b15, b25, &poses);
        poselib_b200::relpose_7pt(b1, b2, &Ms);
        std::vector<Vector3d> b14(b1.begin(), b1.begin() + 4), b24(b2.begin(), b2.begin() + 4);
        poselib_b200::homography_4pt(b14, b24, &H);
        // non-minimal solver and the batch / multi-GPU form
        std::vector<Vector3d> b18(8), b28(8);
        for (int i = 0; i < 8; ++i) {
            b18[i] = {{x1[i](0), x1[i](1), 1.0}};
            b28[i] = {{x2[i](0), x2[i](1), 1.0}};
        }
        Matrix3d E8{};
        poselib_b200::essential_matrix_8pt(b18, b28, &E8);
        poselib_b200::relpose_8pt(b18, b28, &poses);
        std::vector<std::vector<Vector2d>> bx1(3, x1), bx2(3, x2);
        std::vector<Camera> cams(3, cam);
        std::vector<CameraPose> bposes;
        std::vector<std::vector<char>> binl;
        std::vector<RansacStats> bst =
            poselib_b200::estimate_relative_pose_batch<RansacStats>(bx1, bx2, cams, cams, ro, &bposes, &binl, 0, 2);
        RansacStats one = poselib_b200::estimate_relative_pose<RansacStats>(x1, x2, cam, cam, ro, &pose, &inl);
        if (bst.size() != 3 || bst[0].iterations != one.iterations || bst[2].num_inliers != one.num_inliers || binl[1] != inl) {
            std::printf("batch call differs from the single call\n");
            return 1;
        }
        std::printf("adapter run ok\n");
    } else {
        std::printf("adapter link ok (%d devices)\n", plb_device_count());
    }
    return 0;
}
