// A client written against include/poselib_b200.h alone (plain C ABI, no Python, no torch): builds a mixed batch of
// synthetic relative-pose and absolute-pose problems, runs it once on one GPU (plb_ransac_batch) and once over every
// visible GPU in one call from one host thread (plb_ransac_batch_multi), and requires the same results — identical
// iterations, refinements, inlier counts and inlier masks, models to the last bits — problem by problem.  Prints "devices D ok" and exits 0 on success.
// Built by __graft_entry__.build(); run by tests/test_multi_gpu.py on the GPU box.
#include "../include/poselib_b200.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

struct Prob {
    std::vector<double> a, b;
    std::vector<char> inl1, inl2;
};

static void rot_from_axis(const double w[3], double R[9]) {
    const double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double k[3] = {w[0] / th, w[1] / th, w[2] / th}, c = std::cos(th), s = std::sin(th);
    const double K[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
    for (int i = 0; i < 9; ++i) {
        double kk = 0;
        for (int j = 0; j < 3; ++j) kk += K[3 * (i / 3) + j] * K[3 * j + i % 3];
        R[i] = (i % 4 == 0 ? 1.0 : 0.0) + s * K[i] + (1 - c) * kk;
    }
}

int main(int argc, char **argv) {
    const int count = argc > 1 ? std::atoi(argv[1]) : 48;
    const int ndev = plb_device_count();
    if (ndev < 1) {
        std::printf("no device\n");
        return 2;
    }
    std::vector<Prob> P(count);
    std::vector<plb_problem> one(count), multi(count);
    for (int i = 0; i < count; ++i) {
        std::mt19937_64 rng(1000 + i);
        std::uniform_real_distribution<double> U(-1, 1);
        std::normal_distribution<double> G(0, 0.0005);
        const bool rel = (i % 2) == 1;
        const int n = rel ? 1500 + 37 * i : 200 + i;
        double w[3] = {0.3 * U(rng), 0.3 * U(rng), 0.3 * U(rng)}, R[9], t[3] = {U(rng), U(rng), U(rng)};
        rot_from_axis(w, R);
        const double tn = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
        for (double &v : t) v /= tn;
        Prob &p = P[i];
        p.a.resize(2 * (size_t)n);
        p.b.resize((rel ? 2 : 3) * (size_t)n);
        p.inl1.assign(n, 0);
        p.inl2.assign(n, 0);
        for (int k = 0; k < n; ++k) {
            const double u = 0.7 * U(rng), v = 0.7 * U(rng), d = 2.0 + 4.0 * (U(rng) + 1.0);
            const double X[3] = {u * d, v * d, d};
            double Y[3];
            for (int r = 0; r < 3; ++r) Y[r] = R[3 * r] * X[0] + R[3 * r + 1] * X[1] + R[3 * r + 2] * X[2] + t[r];
            const bool outlier = (k % 3) == 0;
            if (rel) {
                p.a[2 * k] = u + G(rng);
                p.a[2 * k + 1] = v + G(rng);
                p.b[2 * k] = outlier ? 0.7 * U(rng) : Y[0] / Y[2] + G(rng);
                p.b[2 * k + 1] = outlier ? 0.7 * U(rng) : Y[1] / Y[2] + G(rng);
            } else {
                p.a[2 * k] = outlier ? 0.7 * U(rng) : Y[0] / Y[2] + G(rng);
                p.a[2 * k + 1] = outlier ? 0.7 * U(rng) : Y[1] / Y[2] + G(rng);
                p.b[3 * k] = X[0];
                p.b[3 * k + 1] = X[1];
                p.b[3 * k + 2] = X[2];
            }
        }
        for (int pass = 0; pass < 2; ++pass) {
            plb_problem &q = pass ? multi[i] : one[i];
            std::memset(&q, 0, sizeof(q));
            q.kind = rel ? PLB_KIND_RELPOSE : PLB_KIND_PNP;
            q.n = (uint64_t)n;
            q.a = p.a.data();
            q.b = p.b.data();
            plb_ransac_opt_default(&q.opt);
            q.opt.max_iterations = 5000;
            q.opt.min_iterations = 100;
            q.opt.seed = (uint64_t)i;
            q.max_error = rel ? 0.002 : 0.01;
            q.model[0] = 1.0;
            q.inliers = pass ? p.inl2.data() : p.inl1.data();
        }
    }
    if (plb_set_device(0) != PLB_OK || plb_ransac_batch(one.data(), (size_t)count, 4) != PLB_OK) {
        std::printf("single-device call failed: %s\n", plb_last_error());
        return 1;
    }
    if (plb_ransac_batch_multi(multi.data(), (size_t)count, 0, 4) != PLB_OK) {
        std::printf("multi-device call failed: %s\n", plb_last_error());
        return 1;
    }
    int bad = 0;
    for (int i = 0; i < count; ++i) {
        const plb_problem &x = one[i], &y = multi[i];
        // Discrete outputs and masks must be identical.  Model coordinates may differ in their last bits: the LM refit of a
        // problem with more than 2048 correspondences is reduced over a thread-block cluster whose size adapts to the
        // number of LO jobs in flight, i.e. to how the batch was split over devices ("Determinism").
        double dm = 0.0;
        for (int k = 0; k < 9; ++k) dm = std::fmax(dm, std::fabs(x.model[k] - y.model[k]));
        const bool same = x.stats.iterations == y.stats.iterations && x.stats.refinements == y.stats.refinements &&
                          x.stats.num_inliers == y.stats.num_inliers &&
                          std::fabs(x.stats.model_score - y.stats.model_score) <= 1e-9 * std::fabs(x.stats.model_score) &&
                          dm <= 2e-5 /* |t| of a relative pose is a gauge of the refiner */ && P[i].inl1 == P[i].inl2 && x.status == PLB_OK && y.status == PLB_OK &&
                          x.stats.num_inliers > 20;
        if (!same) {
            ++bad;
            std::printf("problem %d differs: it %llu/%llu inl %llu/%llu\n", i, (unsigned long long)x.stats.iterations,
                        (unsigned long long)y.stats.iterations, (unsigned long long)x.stats.num_inliers,
                        (unsigned long long)y.stats.num_inliers);
        }
    }
    if (bad) return 1;
    std::printf("devices %d ok\n", ndev);
    return 0;
}
