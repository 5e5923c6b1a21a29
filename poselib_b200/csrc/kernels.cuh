// poselib_b200 — device-side data layout + host-callable launchers (implemented in kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace plb {

enum Kind { KIND_PNP = 0, KIND_RELPOSE = 1, KIND_FUND = 2, KIND_HOMOG = 3 };

// Correspondences of ONE problem, resident in HBM as structure-of-arrays fp64 (exactly the caller's doubles):
//   2D-2D kinds : p[0]=x1.x p[1]=x1.y p[2]=x2.x p[3]=x2.y                      (32 B / correspondence)
//   PnP         : p[0]=x.x  p[1]=x.y  p[2]=X.x  p[3]=X.y  p[4]=X.z             (40 B / correspondence)
// Lanes read consecutive k -> every warp load is one fully coalesced 256 B request per array.
// f[] is the fp32 copy of the same arrays used by the screening pass (16 B / 20 B per correspondence).
struct ProblemDev {
    const double *p[5];
    const float *f[5];
    int n;
    int kind;
    double sq_thr;  // max_error^2 in the units of the points
    int rfc;        // real focal check (fundamental)
};

inline __host__ __device__ int kind_sample_size(int kind) { return kind == KIND_PNP ? 3 : kind == KIND_RELPOSE ? 5 : kind == KIND_FUND ? 7 : 4; }
inline __host__ __device__ int kind_max_models(int kind) { return kind == KIND_PNP ? 4 : kind == KIND_RELPOSE ? 40 : kind == KIND_FUND ? 3 : 1; }
inline __host__ __device__ int kind_model_size(int kind) { return (kind == KIND_PNP || kind == KIND_RELPOSE) ? 7 : 9; }

// Per-round output of the hypothesis kernels.  Models are stored compactly: sample s owns slots
// [first_slot[s], first_slot[s] + n_models[s]) of models / counts / scores; *model_count = total.
// n_models, first_slot, counts and scores may point to pinned host memory (written straight over PCIe).
struct HypOut {
    int *n_models;
    int *first_slot;
    int *model_count;
    uint32_t *counts;
    double *scores;
    double *models;
    float *fscores;   // fast mode: fp32 screening score / count of every model
    uint32_t *fcounts;
};

// LM (local optimisation / final polish) job description — mirrors BundleOptions (types.h:60-95)
struct LmParams {
    int max_iterations;
    int loss_type;   // 0 trivial, 1 truncated, 2 huber, 3 cauchy
    double loss_scale;
    double gradient_tol, step_tol, relative_cost_tol, initial_lambda, min_lambda, max_lambda;
    int subset_mode; // 0: all points; 1: relpose LO subset (Sampson+cheirality inliers at subset_sq_thr of the start pose,
                     //    return untouched if <= 5, estimators/relative_pose.cc:70-76); 2: use given mask
    double subset_sq_thr;
    int use_camera;  // pnp final polish: project with pinhole (fx,fy,cx,cy) instead of the null camera
    double cam[4];
    int score_after; // score the refined model with sq_thr of the problem (count, score)
};
struct LmJobOut {
    double model[9];
    double score;
    uint32_t count;
    int iterations;
    double cost, initial_cost;
};

// ---- launchers (all asynchronous on `stream`) --------------------------------------------------------------
// AoS (caller layout) -> SoA fp64 + fp32.  in_a: 2n doubles; in_b: 2n (2D) or 3n (3D) doubles.
void launch_transpose(const double *in_a, const double *in_b, int n, int b_dim, double *soa64, float *soa32,
                      int n_pad, cudaStream_t stream);
// Fused sample -> solve -> score kernel.  samples: n_samples * K indices.  mode 0 exact, 1 fast (fp32 screen only).
void launch_hypotheses(const ProblemDev &P, const uint32_t *samples, int n_samples, int *work_counter,
                       const HypOut &out, int mode, cudaStream_t stream);
// Exact fp64 scoring of an explicit list of models (model_size doubles each); *n_models_dev == n_models.
void launch_score_models(const ProblemDev &P, const double *models, int n_models, const int *n_models_dev,
                         uint32_t *counts, double *scores, cudaStream_t stream);
// Exact rescoring of selected slots of a HypOut (fast mode confirmation): slots[i] = s*MAXM+m
void launch_rescore_slots(const ProblemDev &P, const HypOut &out, const int *slots, int n_slots, cudaStream_t stream);
// LM refinement: one thread-block cluster per job.  models_in: n_jobs * 9 doubles (model_size used).
// mask (subset_mode 2): n bytes.  idx_scratch: n_jobs * n_pad ints (active-point lists, subset modes 1 and 2).
void launch_lm(const ProblemDev &P, const double *models_in, int n_jobs, const LmParams &prm, const char *mask,
               int *idx_scratch, int n_pad, LmJobOut *out, cudaStream_t stream);
// Final inlier mask of a model (robust/utils.cc:331-351,374-383,434-513)
void launch_inlier_mask(const ProblemDev &P, const double *model, double sq_thr, char *mask, cudaStream_t stream);
// Batched direct solver calls (solvers/*.h surface): one warp per instance.
void launch_solver_batch(int kind, int variant, size_t count, const double *a, const double *b, double *out,
                         int *n_out, int flags, cudaStream_t stream);
int hyp_kernel_blocks(int kind);

} // namespace plb
