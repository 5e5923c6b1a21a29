// ORACLE — TEST INFRASTRUCTURE ONLY.
// The reference's OWN solver benchmark harness (benchmark/solver_benchmark.cc + problem_generator.cc, unmodified, where
// they lie) run for the solvers of the hot path on the mini-Eigen build of the reference's sources.  The harness counts,
// on noise-free random instances from the reference's problem generator, how many returned solutions are VALID
// (is_valid at tol 1e-6) and in how many instances the ground truth is among them (benchmark/solver_benchmark.cc:27-45).
// The reference's main() is renamed and not run (it covers all 40+ solvers); this main() repeats, option for option,
// its configuration of p3p, relpose_5pt, relpose_8pt (8 and 100 points) and homography_4pt (with / without the
// cheirality pre-check) — solver_benchmark.cc:356-360,524-537,600-607 — and prints one line per solver for
// tests/test_reference_unit_tests.py.  (relpose_7pt has no entry in the reference's benchmark.)
#define main reference_benchmark_main
#include "solver_benchmark.cc"
#undef main

#include <cstdio>
#include <cstdlib>

namespace {
void report(const poselib::BenchmarkResult &r) {
    std::printf("%s instances=%d solutions=%d valid=%d found_gt=%d\n", r.name_.c_str(), r.instances_, r.solutions_,
                r.valid_solutions_, r.found_gt_pose_);
}
} // namespace

int main(int argc, char **argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 2000;
    std::srand(1);
    poselib::ProblemOptions options;
    options.camera_fov_ = 75;
    const double tol = 1e-6;

    poselib::ProblemOptions p3p_opt = options;
    p3p_opt.n_point_point_ = 3;
    p3p_opt.n_point_line_ = 0;
    report(poselib::benchmark<poselib::SolverP3P>(n, p3p_opt, tol));
    report(poselib::benchmark<poselib::SolverP3P_lambdatwist>(n, p3p_opt, tol));

    poselib::ProblemOptions rel8pt_opt = options;
    rel8pt_opt.n_point_point_ = 8;
    report(poselib::benchmark_relative<poselib::SolverRel8pt>(n, rel8pt_opt, tol));
    rel8pt_opt.additional_name_ = "(100 pts)";
    rel8pt_opt.n_point_point_ = 100;
    report(poselib::benchmark_relative<poselib::SolverRel8pt>(n, rel8pt_opt, tol));

    poselib::ProblemOptions rel5pt_opt = options;
    rel5pt_opt.n_point_point_ = 5;
    report(poselib::benchmark_relative<poselib::SolverRel5pt>(n, rel5pt_opt, tol));

    poselib::ProblemOptions homo4pt_opt = options;
    homo4pt_opt.n_point_point_ = 4;
    report(poselib::benchmark_homography<poselib::SolverHomography4pt<false>>(n, homo4pt_opt, tol));
    report(poselib::benchmark_homography<poselib::SolverHomography4pt<true>>(n, homo4pt_opt, tol));
    return 0;
}
