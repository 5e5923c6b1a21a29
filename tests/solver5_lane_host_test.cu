// CPU-side check of poselib_b200/csrc/solver5_lane.cuh: the thread-per-sample first half of relpose_5pt is
// __host__ __device__, so the SAME source k5_prep_lane runs on the device is compiled for the host here (nvcc host pass;
// the host compiler is asked not to contract) and its outputs are printed as hex doubles;
// tests/test_solver5_lane_host.py compares them bit for bit with the oracle's intermediates.
#include "../poselib_b200/csrc/solver5_lane.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

int main() {
    // stdin: count, then per sample 30 doubles (x1s 15 | x2s 15) as hex floats
    int count = 0;
    if (std::scanf("%d", &count) != 1) return 2;
    for (int s = 0; s < count; ++s) {
        double xs[30];
        for (int i = 0; i < 30; ++i) {
            char tok[64];
            if (std::scanf("%63s", tok) != 1) return 2;
            xs[i] = std::strtod(tok, nullptr);
        }
        std::vector<double> W(256);  // 100 left block | 10 scratch | 100 parked right-hand sides
        double Nb[36], A[39], cp[11];
        plb::lane5::solve_5pt_poly_lane(W.data(), xs, Nb, A, cp);
        for (int i = 0; i < 36; ++i) std::printf("%a ", Nb[i]);
        for (int i = 0; i < 39; ++i) std::printf("%a ", A[i]);
        for (int i = 0; i < 11; ++i) std::printf("%a ", cp[i]);
        std::printf("\n");
    }
    return 0;
}
