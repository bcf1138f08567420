/*
 * bo_b200.h -- C ABI of the B200-native GP surrogate + acquisition hot path.
 *
 * Drop-in boundary for billbearhunter/BayesianOptimizer's surrogate/acquisition seam.  The
 * reference has no FFI of its own: the path sits behind Python calls into botorch/gpytorch
 * (SURVEY.md section 8b).  Each entry point below names the reference call site it replaces.
 * The Python host (bayesianoptimizer_b200/engine.py) binds these with ctypes and passes
 * `tensor.data_ptr()` of contiguous float64 CUDA tensors plus the current CUDA stream.
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross the ABI; every call returns a status.
 *   - status 0 = ok; status k > 0 (bo_fit, bo_append only) = matrix not positive definite at
 *     1-based pivot k (the caller retries with a larger jitter, as optimization/Bayesian6.py:482-488
 *     does); status < 0 = one of the BO_E_* codes, text via bo_last_error().
 *   - "_dev" pointers are device pointers on the handle's device, borrowed for the call only;
 *     "_host" pointers are host pointers.  All matrices are row-major float64.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous with respect to the host unless stated otherwise.
 *   - the handle owns all device workspaces (L, packed L^-1, alpha, scaled X, sweep panels) and
 *     frees them in bo_destroy(); bo_release_workspace() drops the sweep scratch early so that a
 *     co-resident simulator (simulation/taichi.py:10) gets its HBM back.
 *   - there is no CPU fallback: every compute entry point fails with BO_E_CUDA if no sm_100 device
 *     is usable.
 */
#ifndef BO_B200_H
#define BO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define BO_ABI_VERSION 1

/* kernel kinds: gpytorch MaternKernel(nu=2.5, ard) (optimization/Bayesian6.py:471-473,
 * optimization/Bayesian7.py:162-166) and RBFKernel(ard) (botorch>=0.12 SingleTaskGP default,
 * optimization/Bayesian.py:91), both under an (optional) ScaleKernel outputscale. */
#define BO_KERNEL_MATERN52 0
#define BO_KERNEL_RBF      1
#define BO_KERNEL_LINEAR_MATERN52 2   /* ScaleKernel(LinearKernel + MaternKernel(2.5, ard)): outputscale * (v <x,x'> + matern),
                                       * the explicit kernel of optimization/Bayesian6.py:471-473 and Bayesian7.py:162-166;
                                       * supported by every entry point (fit / posterior / sweep / append / lml / acq_grad / refine) */

/* acquisition kinds (analytic closed forms; maximisation as in optimization/Bayesian.py:98) */
#define BO_ACQ_EI    0   /* expected improvement                                  (Bayesian.py:100-101) */
#define BO_ACQ_LOGEI 1   /* log expected improvement, erfcx tails                 (Bayesian.py:9,101)   */
#define BO_ACQ_UCB   2   /* mean + sqrt(beta) * sigma                                                   */
#define BO_ACQ_VAR   3   /* posterior variance score (active-learning sweep,      Bayesian7.py:670-671) */
#define BO_ACQ_MEAN  4   /* posterior mean                                                              */

/* contraction used by bo_sweep for the variance term u = L^-1 k* (bo_set_sweep_mode) */
#define BO_SWEEP_AUTO 0   /* INT8-sliced tensor path (BO_SWEEP_I8X7 up to 16 384 padded observations, I8X8 above; guarded) for large
                           * pools of an exact GP, FP64 DMMA otherwise (see bo_set_sweep_mode) */
#define BO_SWEEP_FP64 1   /* always the FP64 DMMA contraction */
#define BO_SWEEP_I8X7 2   /* INT8-sliced, 7 slices per operand: one 7-bit + six 8-bit digits (54-bit operands, 28 slice products) */
#define BO_SWEEP_I8X8 3   /* INT8-sliced, 8 slices per operand: eight 7-bit digits (55-bit operands, 36 slice products) */

/* error codes (negative statuses) */
#define BO_E_INVALID  (-1)  /* bad argument                                  */
#define BO_E_CUDA     (-2)  /* CUDA runtime error / no usable device         */
#define BO_E_NOMEM    (-3)  /* device allocation failed                      */
#define BO_E_NOTFIT   (-4)  /* posterior/sweep/refine/append before bo_fit   */
#define BO_E_CAPACITY (-5)  /* n or d or topk beyond the compiled limits     */

#define BO_MAX_DIM   16     /* input dimension limit (reference: d = 5)      */
#define BO_MAX_SELECT 8192   /* largest K of bo_topk_scores (K_BIG_CAP = 8000, optimization/Bayesian7.py:66) */
#define BO_MAX_TOPK  64     /* in-kernel top-k limit                         */
#define BO_SOBOL_BITS 30    /* torch.quasirandom.SobolEngine.MAXBIT          */

typedef struct bo_handle bo_handle;

/* Scrambled-Sobol generator state: direction numbers and digital shift of a fresh
 * torch.quasirandom.SobolEngine(d, scramble=True, seed) (`.sobolstate`, `.shift`).  Candidate i of
 * the pool is generated inside the sweep kernel from its GLOBAL index, bit-identical to row i of
 * SobolEngine.draw(dtype=float64); no candidate bytes cross HBM.  Replaces the host-side pool
 * construction of optimization/Bayesian7.py:650-655 and optimize_acqf's raw samples
 * (optimization/Bayesian.py:105-112). */
typedef struct bo_sobol {
    int32_t  d;
    uint32_t direction[BO_MAX_DIM][BO_SOBOL_BITS];
    uint32_t shift[BO_MAX_DIM];
} bo_sobol;

/* library / device ------------------------------------------------------------------------- */
int         bo_abi_version(void);
int         bo_device_count(void);
int         bo_create(bo_handle** out, int device);
void        bo_destroy(bo_handle* h);
const char* bo_last_error(const bo_handle* h);
int         bo_release_workspace(bo_handle* h);

/* GP fit: K = k(X,X) + (noise + jitter) I, L = chol(K), alpha = K^-1 (y - mean), explicit L^-1.
 * Replaces SingleTaskGP(train_X, train_Y) + the prediction caches built inside model.posterior
 * (optimization/Bayesian.py:89-94; SURVEY.md 3.3: mean_cache -> alpha, covar_cache -> L^-1).
 * X_dev[n,d], y_dev[n]; lengthscale_host[d].  Synchronises the stream (the pivot status is read back). */
int bo_fit(bo_handle* h, const double* X_dev, const double* y_dev, int32_t n, int32_t d,
           int32_t kernel_kind, const double* lengthscale_host, double outputscale, double noise,
           double mean, double jitter, void* stream);

/* Same with HOST X/y (copies inside) -- the entry the end-to-end benchmark times. */
int bo_fit_host(bo_handle* h, const double* X_host, const double* y_host, int32_t n, int32_t d,
                int32_t kernel_kind, const double* lengthscale_host, double outputscale, double noise,
                double mean, double jitter, void* stream);

/* General form of bo_fit / bo_fit_host: adds the LinearKernel variance v of BO_KERNEL_LINEAR_MATERN52 (ignored for the
 * stationary kinds) and takes host or device X / y (host_inputs != 0: host pointers, copied inside). */
int bo_fit_ex(bo_handle* h, const double* X, const double* y, int32_t n, int32_t d, int32_t kernel_kind,
              const double* lengthscale_host, double outputscale, double noise, double mean, double jitter,
              double linear_variance, int32_t host_inputs, void* stream);

/* Per-dimension LinearKernel variances v[d] for the NEXT bo_fit_ex / bo_svgp_load of the BO_KERNEL_LINEAR_MATERN52 kind:
 * k = outputscale * (sum_k v_k x_k x'_k + matern).  gpytorch's LinearKernel(ard_num_dims = D, batch_shape = [T]) of
 * optimization/Bayesian7.py:162-166 registers raw_variance with shape (T, 1, D), one variance per input dimension; the
 * scalar `linear_variance` argument of those calls covers the (T, 1, 1) layout.  Consumed by that fit; NULL / d = 0 clears.
 * The batched LML (bo_lml_grad_batched) keeps one scalar variance as its hyper-parameter. */
int bo_set_linear_variance_ard(bo_handle* h, const double* variances_host, int32_t d);

/* SVGP predictive state: loads ONE task of the reference's batched sparse variational GP
 * (optimization/Bayesian7.py:129-195: whitened VariationalStrategy, CholeskyVariationalDistribution, kernel
 * ScaleKernel(Linear + Matern-5/2), ConstantMean, GaussianLikelihood) into the handle, after which bo_posterior / bo_sweep
 * evaluate its predictive distribution exactly as the chunked pool scan of Bayesian7.py:664-671 does:
 *   L = chol(k(Z,Z) + jitter I),  u = L^-1 k(Z,x*),  mean = mean_const + u^T m,
 *   var = k(x*,x*) + jitter - ||u||^2 + ||Ls^T u||^2 + noise          (S = Ls Ls^T, clamped at min_variance)
 * Z_dev[M,d] inducing points (in the model's own input space), var_mean_dev[M] = m, var_chol_dev[M,M] = Ls (lower,
 * row-major; the upper triangle is ignored), jitter = gpytorch's variational_cholesky_jitter (1e-6 double / 1e-4 float),
 * noise = the task's likelihood noise (0 for the latent variance).  Returns 0, a negative error or the failing pivot.
 * bo_append / bo_refine / bo_acq_grad / bo_posterior_multi refuse a handle in this state; bo_fit returns it to the exact mode. */
int bo_svgp_load(bo_handle* h, const double* Z_dev, int32_t M, int32_t d, int32_t kernel_kind, const double* lengthscale_host,
                 double outputscale, double linear_variance, double mean, double noise, double jitter,
                 const double* var_mean_dev, const double* var_chol_dev, void* stream);

/* introspection of the fitted state (device outputs; any may be NULL):
 *   alpha_dev[n], chol_dev[n,n] (lower triangle, upper zero), linv_dev[n,n] (lower, upper zero) */
int bo_get_state(bo_handle* h, double* alpha_dev, double* chol_dev, double* linv_dev, void* stream);
int bo_num_obs(const bo_handle* h);

/* Posterior mean / variance at Xs_dev[N,d]: mean = m + k*^T alpha ; var = max(s2 - ||L^-1 k*||^2,
 * min_variance).  Replaces model.posterior(X).mean/.variance
 * (optimization/Bayesian2.py:168-171, optimization/Bayesian6.py:615-617, Bayesian7.py:666-671). */
int bo_posterior(bo_handle* h, const double* Xs_dev, int64_t N, double min_variance,
                 double* mean_dev, double* var_dev, void* stream);

/* m outputs that share the fitted kernel matrix (one Cholesky, m right-hand sides): posterior means mean_dev[N,m] at
 * Xs_dev[N,d] for targets Y_dev[n,m] with constant means means_host[m] (NULL = zeros); the shared variance goes to
 * var_dev[N] when non-NULL.  Exact-GP counterpart of the batched 8-task models and of model.posterior(X).mean with
 * (N, m) outputs (optimization/Bayesian1.py:109-113, Bayesian2.py:168-171, Bayesian6.py:615-617, Bayesian7.py:129-195). */
int bo_posterior_multi(bo_handle* h, const double* Y_dev, int32_t m, const double* means_host, const double* Xs_dev,
                       int64_t N, double min_variance, double* mean_dev, double* var_dev, void* stream);

/* Acquisition sweep over candidates [first_index, first_index + N) of a pool:
 *   cand_dev != NULL : explicit pool, cand_dev[N,d] holds exactly this shard's rows;
 *   cand_dev == NULL : in-kernel scrambled Sobol from `sobol_host` keyed by the global index.
 * Fused on device: K(X*,X) panel -> mean -> variance (contraction with packed L^-1: FP64 DMMA, or exact INT8 slice
 * products on tcgen05 for large pools -- bo_set_sweep_mode) ->
 * EI/LogEI/UCB -> top-k by (value desc, global index asc).  Outputs (device): vals_dev[topk],
 * idx_dev[topk] (entries beyond N are -inf / -1); optional per-candidate mean_dev/var_dev/acq_dev[N].
 * Replaces the chunked pool scan + CPU topk of optimization/Bayesian7.py:664-682 and the
 * raw-sample scoring inside optimize_acqf (optimization/Bayesian.py:105-112). */
int bo_sweep(bo_handle* h, int32_t acq_kind, double best_f, double beta, double min_variance,
             const double* cand_dev, const bo_sobol* sobol_host, int64_t first_index, int64_t N,
             int32_t topk, double* vals_dev, int64_t* idx_dev,
             double* mean_dev, double* var_dev, double* acq_dev, void* stream);

/* Same with HOST outputs vals_host[topk], idx_host[topk] (and host candidates when cand_host != NULL);
 * synchronises.  The end-to-end entry: host buffers in, host buffers out. */
int bo_sweep_host(bo_handle* h, int32_t acq_kind, double best_f, double beta, double min_variance,
                  const double* cand_host, const bo_sobol* sobol_host, int64_t first_index, int64_t N,
                  int32_t topk, double* vals_host, int64_t* idx_host, void* stream);

/* Materialise pool points [first_index, first_index+N) of the Sobol stream into out_dev[N,d]
 * (used to turn winning indices back into coordinates; Bayesian7.py:682 `cand_unit_cpu[idxs_big]`). */
int bo_sobol_points(bo_handle* h, const bo_sobol* sobol_host, const int64_t* idx_dev, int64_t N,
                    double* out_dev, void* stream);

/* Batched projected-gradient-ascent refinement of k starts inside [0,1]^d with analytic gradients
 * of the acquisition; replaces gen_candidates_scipy (L-BFGS-B, autograd) inside optimize_acqf
 * (optimization/Bayesian.py:105-112).  starts_dev[k,d] -> x_dev[k,d], val_dev[k]. */
int bo_refine(bo_handle* h, int32_t acq_kind, double best_f, double beta, double min_variance,
              const double* starts_dev, int32_t k, int32_t iters, double* x_dev, double* val_dev,
              void* stream);

/* Acquisition value and gradient at Xq_dev[k,d] (building block of bo_refine; the autograd
 * backward of acq(X) in the reference). val_dev[k], grad_dev[k,d]. */
int bo_acq_grad(bo_handle* h, int32_t acq_kind, double best_f, double beta, double min_variance,
                const double* Xq_dev, int32_t k, double* val_dev, double* grad_dev, void* stream);

/* Kriging-believer / observation append: border L and L^-1 with one row.
 * use_believer != 0 appends y = mu(x) (alpha' = [alpha; 0]); otherwise appends the observed y.
 * Replaces the refit that the torch.cat register step triggers (optimization/Bayesian.py:143-148)
 * and the greedy set_X_pending loop (optimization/Bayesian6.py:908-919). x_dev[d]. */
int bo_append(bo_handle* h, const double* x_dev, double y, int32_t use_believer, void* stream);

/* Batched exact log marginal likelihood and gradient over R hyper-parameter restarts on the
 * given X/y.  theta_host[R, p] = log lengthscale[d], log outputscale, log noise (p = d+2), plus log linear variance
 * for BO_KERNEL_LINEAR_MATERN52 (p = d+3).  lml_host[R]; grad_host[R, p] (w.r.t. the log parameters);
 * status_host[R] (0 or pivot).
 * Replaces the ExactMarginalLogLikelihood closure fit_gpytorch_mll evaluates
 * (optimization/Bayesian.py:92-93, optimization/Bayesian6.py:480-488). */
int bo_lml_grad_batched(bo_handle* h, const double* X_dev, const double* y_dev, int32_t n, int32_t d,
                        int32_t kernel_kind, double mean, const double* theta_host, int32_t R,
                        double* lml_host, double* grad_host, int32_t* status_host, void* stream);

/* Large top-K of a dense score array: the K <= BO_MAX_SELECT best entries of scores_dev[N] in the order of the fused
 * sweep's top-k (value descending, global index = first_index + position ascending, NaN last) to vals_dev[K] /
 * idx_dev[K]; entries past N are (-inf, -1).  Radix select + ordered tie compaction + one-CTA bitonic sort.
 * Replaces `torch.topk(unc, K_big)` on the CPU copy of all scores (optimization/Bayesian7.py:671-682). */
int bo_topk_scores(bo_handle* h, const double* scores_dev, int64_t N, int64_t first_index, int32_t K, double* vals_dev,
                   int64_t* idx_dev, void* stream);

/* Greedy farthest-point sampling of m of the N points X_dev[N,d], starting from index `start`; picks go to
 * idx_dev[m] in selection order (arg-max of the running min squared distance, first index on ties).
 * Replaces farthest_point_sampling (optimization/Bayesian7.py:82-107, Bayesian6.py:88-107), which the reference
 * runs on the CPU over the top-K pool (Bayesian7.py:679-688) and for inducing-point selection. */
int bo_fps(bo_handle* h, const double* X_dev, int64_t N, int32_t d, int32_t m, int64_t start, int64_t* idx_dev,
           void* stream);

/* FP64 peak probe (DMMA.8x8x4 register-resident loop): the roofline denominator that
 * MEASURED_PEAKS.json lacks.  Returns TFLOP/s in *tflops_host. */
int bo_fp64_peak(bo_handle* h, int32_t use_dmma, double seconds, double* tflops_host);

/* Throughput probe of the fit's grouped FP64 DMMA GEMM (C = A B^T + C, one m x n x k problem; cfg 0: 64x64 tiles,
 * 1: 128x128, 2: 128x64).  Returns TFLOP/s in *tflops_host.  Diagnostic only. */
int bo_gemm_probe(bo_handle* h, int32_t m, int32_t n, int32_t k, int32_t cfg, int32_t reps, double* tflops_host);

/* Choose how bo_sweep contracts L^-1 with the K(X, X*) panel (BO_SWEEP_*; default BO_SWEEP_AUTO).  The INT8-sliced path
 * computes the same FP64 quantity: both operands are cut into S signed int8 slices (error-free, Ozaki scheme I), the
 * S (S + 1) / 2 leading slice products run on the INT8 tensor cores (tcgen05.mma kind::i8) with exact INT32
 * accumulation and are recombined in FP64.  Two digit geometries: I8X8 = eight 7-bit digits (55-bit operands, 36 products),
 * I8X7 = one 7-bit + six 8-bit digits (54-bit operands, 28 products: 3.3x the slicing error of I8X8 on the reference's data,
 * 22 % fewer MMAs; INT32 accumulators hold up to 16 384 padded observations, above that I8X7 resolves to I8X8).
 * Every sliced sweep carries a per-candidate accuracy guard: the slicing error of ||u||^2 is bounded from the row scales of
 * L^-1 and ||u||^2 itself, and a candidate whose variance is not at least 5e8 x that bound (slicing error <= 2e-9 relative)
 * is re-scored on the FP64 DMMA contraction inside the same call -- same outputs, same top-k order;
 * bo_last_sweep_flagged() reports how many (candidates with sigma^2 below ~2e-4 (I8X8) / ~7e-4 (I8X7) of the prior variance:
 * next to training rows).  Eligible models: exact GP of any kernel kind (the linear + Matern kind scales each candidate's
 * operand by its own bound on |k*|) with at least 256 (padded) observations, and SVGP predictive states (bo_svgp_load): there
 * the sliced sweep contracts the stacked factor [L^-1; Ls^T L^-1] with the same sliced panel in ONE pass -- the second variance
 * term ||Ls^T u||^2 = ||(Ls^T L^-1) k*||^2 -- with the guard covering both terms (the FP64 form runs two triangular passes).
 * The pinned modes depend on the model only, so all shards of a pool take the same path (bit-identical values for
 * every shard layout); AUTO additionally keeps pools below 2 x SMs x 64 candidates (SVGP states: below SMs / 2 x 64, they have
 * no row-split FP64 kernel) and models with fewer than 512 (padded) observations on the FP64 path.  bo_posterior / bo_posterior_multi stay on the FP64 contraction under AUTO
 * (model.posterior numerics must not depend on N) and follow a pinned mode.  A sliced sweep synchronises the stream once
 * (the count of flagged candidates is read back). */
int bo_set_sweep_mode(bo_handle* h, int32_t mode);

/* The pinned mode (BO_SWEEP_FP64 / I8X7 / I8X8) the handle's current mode resolves to for a pool of pool_total
 * candidates of the fitted model: callers that shard one pool over several calls or GPUs resolve once on the global
 * size and pin the result, so that every shard takes the same path. */
int bo_resolve_sweep_mode(const bo_handle* h, int64_t pool_total);

/* Contraction the last bo_sweep ran: 0 = FP64 DMMA, 7 / 8 = INT8-sliced with that many slices, -1 = no sweep yet. */
int bo_last_sweep_path(const bo_handle* h);

/* Candidates of the last sliced sweep that its accuracy guard sent through the FP64 contraction (0 for an FP64 sweep).
 * Which path a candidate takes depends on the candidate alone, never on the pool or shard it arrives in. */
int64_t bo_last_sweep_flagged(const bo_handle* h);

/* INT8 tensor-pipe peak probe (tcgen05.mma kind::i8, 128 x 256 x 32 on resident operands): roofline denominator
 * of the sliced sweep.  Returns TOP/s (multiply and add counted separately) in *tops_host. */
int bo_i8_peak(bo_handle* h, double seconds, double* tops_host);

/* Kernel-launch counter (own kernels launched through this handle since creation). */
int64_t bo_launch_count(const bo_handle* h);

/* Device time of the last bo_sweep's fused kernel in milliseconds (CUDA events on the sweep's
 * stream; synchronises on the stop event). Negative if no sweep has run. */
double bo_last_sweep_ms(bo_handle* h);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* BO_B200_H */
