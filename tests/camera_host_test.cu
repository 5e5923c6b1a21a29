// CPU-side check of poselib_b200/csrc/camera.cuh: the camera functions are __host__ __device__, so the SAME source the
// layout kernel runs on the device is compiled for the host here (nvcc host pass, -fmad=false has no host effect; the
// host compiler is asked not to contract) and printed as hex doubles; tests/test_camera_host_build.py compares them
// bit for bit with the oracle's camera models.
#include "../poselib_b200/csrc/camera.cuh"
#include <cstdio>
#include <cstdlib>

using namespace plb;

int main(int argc, char **argv) {
    // argv: model p0..p7 then pairs u v
    if (argc < 10) return 2;
    CamDev c;
    c.model = std::atoi(argv[1]);
    c.reserved = 0;
    for (int i = 0; i < 8; ++i) c.p[i] = std::strtod(argv[2 + i], nullptr);
    for (int a = 10; a + 1 < argc; a += 2) {
        const double u = std::strtod(argv[a], nullptr), v = std::strtod(argv[a + 1], nullptr);
        double d[3], M[6], ox, oy;
        cam_unproject_with_jac(c, u, v, d, M);
        cam_unproject2(c, u, v, ox, oy);
        double pu, pv, J[6], qu, qv;
        cam_project_with_jac(c, d[0], d[1], d[2], pu, pv, J);
        cam_project(c, d[0], d[1], d[2], qu, qv);
        std::printf("%a %a %a", d[0], d[1], d[2]);
        for (int i = 0; i < 6; ++i) std::printf(" %a", M[i]);
        std::printf(" %a %a %a %a", ox, oy, pu, pv);
        for (int i = 0; i < 6; ++i) std::printf(" %a", J[i]);
        std::printf(" %a %a\n", qu, qv);
    }
    return 0;
}
