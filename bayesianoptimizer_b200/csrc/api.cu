// C ABI (include/bo_b200.h) of the B200 GP surrogate + acquisition library.
#include "gemm.cuh"
#include <new>

using namespace bo;

namespace bo {
int acq_grad_impl(bo_handle* h, int acq_kind, double best_f, double beta, double min_var,
                  const double* Xq_dev, int k, double* val_dev, double* grad_dev, cudaStream_t st);
int refine_impl(bo_handle* h, int acq_kind, double best_f, double beta, double min_var,
                const double* starts_dev, int k, int iters, double* x_dev, double* val_dev, cudaStream_t st);
int append_impl(bo_handle* h, const double* x_dev, double y, int use_believer, cudaStream_t st);
int posterior_multi_impl(bo_handle* h, const double* Y_dev, int m, const double* means_host, const double* Xs_dev, int64_t N,
                         double min_var, double* mean_dev, double* var_dev, cudaStream_t st);
int topk_scores_impl(bo_handle* h, const double* scores_dev, int64_t N, int64_t first_index, int K, double* vals_dev,
                     int64_t* idx_dev, cudaStream_t st);
int fps_impl(bo_handle* h, const double* X_dev, int64_t N, int d, int m, int64_t start, int64_t* idx_dev, cudaStream_t st);
int gemm_probe_impl(bo_handle* h, int m, int n, int k, int cfg, int reps, double* tflops);
int export_state(bo_handle* h, double* alpha_dev, double* chol_dev, double* linv_dev, cudaStream_t st);
int lml_impl(bo_handle* h, const double* X_dev, const double* y_dev, int n, int d, int kind, double mean,
             const double* theta_host, int R, double* lml_host, double* grad_host, int* status_host,
             cudaStream_t st);
}

namespace bo {
int create_handle(bo_handle** out, int device) {
    if (!out) return BO_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return BO_E_CUDA; }
    if (device < 0 || device >= ndev) return BO_E_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return BO_E_CUDA; }
    if (prop.major != 10) return BO_E_CUDA;      // built for sm_100a only; no fallback
    bo_handle* h = new (std::nothrow) bo_handle();
    if (!h) return BO_E_NOMEM;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaMalloc(&h->info_dev, sizeof(int)) != cudaSuccess ||
        cudaMallocHost(&h->info_host, sizeof(int)) != cudaSuccess ||
        cudaMalloc(&h->out_stage_val, BO_MAX_TOPK * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->out_stage_idx, BO_MAX_TOPK * sizeof(int64_t)) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
        gemm_init(h) != 0) {
        cudaGetLastError();
        bo_destroy(h);
        return BO_E_CUDA;
    }
    *out = h;
    return 0;
}
}  // namespace bo

extern "C" {

int bo_abi_version(void) { return BO_ABI_VERSION; }

int bo_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int bo_create(bo_handle** out, int device) { return create_handle(out, device); }

int bo_release_workspace(bo_handle* h) {
    if (!h) return BO_E_INVALID;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->panel) cudaFree(h->panel);
    h->panel = nullptr; h->panel_bytes = 0;
    if (h->panel2) cudaFree(h->panel2);
    h->panel2 = nullptr; h->panel2_bytes = 0;
    if (h->cand_stage) cudaFree(h->cand_stage);
    h->cand_stage = nullptr; h->cand_stage_bytes = 0;
    if (h->panel8) cudaFree(h->panel8);
    h->panel8 = nullptr; h->panel8_bytes = 0;
    if (h->Lp8) cudaFree(h->Lp8);                 // re-sliced from L^-1 by the next sliced sweep
    h->Lp8 = nullptr; h->Lp8_bytes = 0; h->Lp8_epoch = 0;
    if (h->rowscale) cudaFree(h->rowscale);
    h->rowscale = nullptr; h->rowscale_cap = 0;
    if (h->flag_idx) cudaFree(h->flag_idx);
    h->flag_idx = nullptr; h->flag_cap = 0;
    lml_release(h);
    return 0;
}

void bo_destroy(bo_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    bo_release_workspace(h);
    lml_release(h);
    void* ptrs[] = {h->qbuf, h->split_ws, h->Xs, h->Xraw, h->yv, h->alpha, h->Lm, h->Li, h->Tw, h->Lp, h->vec1, h->vec2, h->vec3,
                    h->info_dev, h->plan_dev, h->part_val, h->part_idx, h->sobol_dev,
                    h->out_stage_val, h->out_stage_idx, h->Lp2, h->select_ws, h->Lp8, h->rowscale, h->guard_dev, h->flag_count_dev, h->svB};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->info_host) cudaFreeHost(h->info_host);
    if (h->flag_count_host) cudaFreeHost(h->flag_count_host);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    cudaGetLastError();
    delete h;
}

const char* bo_last_error(const bo_handle* h) { return h ? h->err.c_str() : "null handle"; }
int bo_num_obs(const bo_handle* h) { return (h && h->fitted) ? h->n : 0; }
int64_t bo_launch_count(const bo_handle* h) { return h ? h->launches : 0; }

int bo_fit_ex(bo_handle* h, const double* X, const double* y, int32_t n, int32_t d, int32_t kernel_kind,
              const double* lengthscale_host, double outputscale, double noise, double mean, double jitter,
              double linear_variance, int32_t host_inputs, void* stream) {
    if (!h) return BO_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (!host_inputs)
        return fit_impl(h, X, y, n, d, kernel_kind, lengthscale_host, outputscale, noise, mean, jitter, linear_variance, st);
    if (n < 1 || d < 1 || !X || !y) return fail(h, BO_E_INVALID, "bo_fit_host: bad argument");
    BO_CUDA(h, cudaSetDevice(h->device));
    const size_t need = ((size_t)n * d + n) * sizeof(double);
    if (need > h->cand_stage_bytes) {
        if (h->cand_stage) cudaFree(h->cand_stage);
        h->cand_stage = nullptr; h->cand_stage_bytes = 0;
        BO_CUDA(h, cudaMalloc(&h->cand_stage, need));
        h->cand_stage_bytes = need;
    }
    double* Xd = h->cand_stage; double* yd = Xd + (size_t)n * d;
    BO_CUDA(h, cudaMemcpyAsync(Xd, X, (size_t)n * d * 8, cudaMemcpyHostToDevice, st));
    BO_CUDA(h, cudaMemcpyAsync(yd, y, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    return fit_impl(h, Xd, yd, n, d, kernel_kind, lengthscale_host, outputscale, noise, mean, jitter, linear_variance, st);
}

int bo_fit(bo_handle* h, const double* X_dev, const double* y_dev, int32_t n, int32_t d, int32_t kernel_kind,
           const double* lengthscale_host, double outputscale, double noise, double mean, double jitter,
           void* stream) {
    return bo_fit_ex(h, X_dev, y_dev, n, d, kernel_kind, lengthscale_host, outputscale, noise, mean, jitter, 0.0, 0, stream);
}

int bo_fit_host(bo_handle* h, const double* X_host, const double* y_host, int32_t n, int32_t d,
                int32_t kernel_kind, const double* lengthscale_host, double outputscale, double noise,
                double mean, double jitter, void* stream) {
    return bo_fit_ex(h, X_host, y_host, n, d, kernel_kind, lengthscale_host, outputscale, noise, mean, jitter, 0.0, 1, stream);
}

int bo_topk_scores(bo_handle* h, const double* scores_dev, int64_t N, int64_t first_index, int32_t K, double* vals_dev,
                   int64_t* idx_dev, void* stream) {
    if (!h) return BO_E_INVALID;
    return topk_scores_impl(h, scores_dev, N, first_index, K, vals_dev, idx_dev, (cudaStream_t)stream);
}

int bo_svgp_load(bo_handle* h, const double* Z_dev, int32_t M, int32_t d, int32_t kernel_kind, const double* lengthscale_host,
                 double outputscale, double linear_variance, double mean, double noise, double jitter,
                 const double* var_mean_dev, const double* var_chol_dev, void* stream) {
    if (!h) return BO_E_INVALID;
    return svgp_load_impl(h, Z_dev, M, d, kernel_kind, lengthscale_host, outputscale, linear_variance, mean, noise, jitter,
                          var_mean_dev, var_chol_dev, (cudaStream_t)stream);
}

int bo_get_state(bo_handle* h, double* alpha_dev, double* chol_dev, double* linv_dev, void* stream) {
    if (!h) return BO_E_INVALID;
    if (!h->fitted) return fail(h, BO_E_NOTFIT, "bo_get_state before a successful bo_fit");
    return export_state(h, alpha_dev, chol_dev, linv_dev, (cudaStream_t)stream);
}

int bo_posterior(bo_handle* h, const double* Xs_dev, int64_t N, double min_variance, double* mean_dev,
                 double* var_dev, void* stream) {
    if (!h) return BO_E_INVALID;
    if (!Xs_dev && N > 0) return fail(h, BO_E_INVALID, "bo_posterior: null candidates");
    // model.posterior numerics do not depend on N: AUTO keeps bo_posterior on the FP64 contraction, the sliced path is opt-in
    // here (a pinned BO_SWEEP_I8X* mode)
    return sweep_impl(h, BO_ACQ_MEAN, 0.0, 0.0, min_variance, Xs_dev, nullptr, 0, N, 0, nullptr, nullptr,
                      mean_dev, var_dev, nullptr, (cudaStream_t)stream, h->sweep_mode == BO_SWEEP_AUTO ? BO_SWEEP_FP64 : -1);
}

int bo_posterior_multi(bo_handle* h, const double* Y_dev, int32_t m, const double* means_host, const double* Xs_dev,
                       int64_t N, double min_variance, double* mean_dev, double* var_dev, void* stream) {
    if (!h) return BO_E_INVALID;
    return posterior_multi_impl(h, Y_dev, m, means_host, Xs_dev, N, min_variance, mean_dev, var_dev, (cudaStream_t)stream);
}

int bo_sweep(bo_handle* h, int32_t acq_kind, double best_f, double beta, double min_variance,
             const double* cand_dev, const bo_sobol* sobol_host, int64_t first_index, int64_t N, int32_t topk,
             double* vals_dev, int64_t* idx_dev, double* mean_dev, double* var_dev, double* acq_dev, void* stream) {
    if (!h) return BO_E_INVALID;
    return sweep_impl(h, acq_kind, best_f, beta, min_variance, cand_dev, sobol_host, first_index, N, topk,
                      vals_dev, idx_dev, mean_dev, var_dev, acq_dev, (cudaStream_t)stream);
}

int bo_sweep_host(bo_handle* h, int32_t acq_kind, double best_f, double beta, double min_variance,
                  const double* cand_host, const bo_sobol* sobol_host, int64_t first_index, int64_t N,
                  int32_t topk, double* vals_host, int64_t* idx_host, void* stream) {
    if (!h) return BO_E_INVALID;
    if (topk < 1 || topk > BO_MAX_TOPK || !vals_host || !idx_host) return fail(h, BO_E_INVALID, "bo_sweep_host: bad topk/outputs");
    if (!h->fitted) return fail(h, BO_E_NOTFIT, "sweep before a successful bo_fit");
    cudaStream_t st = (cudaStream_t)stream;
    BO_CUDA(h, cudaSetDevice(h->device));
    const double* cand_dev = nullptr;
    if (cand_host) {
        const size_t need = (size_t)N * h->d * sizeof(double);
        if (need > h->cand_stage_bytes) {
            if (h->cand_stage) cudaFree(h->cand_stage);
            h->cand_stage = nullptr; h->cand_stage_bytes = 0;
            BO_CUDA(h, cudaMalloc(&h->cand_stage, need ? need : 8));
            h->cand_stage_bytes = need ? need : 8;
        }
        if (need) BO_CUDA(h, cudaMemcpyAsync(h->cand_stage, cand_host, need, cudaMemcpyHostToDevice, st));
        cand_dev = h->cand_stage;
    }
    int rc = sweep_impl(h, acq_kind, best_f, beta, min_variance, cand_dev, sobol_host, first_index, N, topk,
                        h->out_stage_val, h->out_stage_idx, nullptr, nullptr, nullptr, st);
    if (rc) return rc;
    BO_CUDA(h, cudaMemcpyAsync(vals_host, h->out_stage_val, topk * sizeof(double), cudaMemcpyDeviceToHost, st));
    BO_CUDA(h, cudaMemcpyAsync(idx_host, h->out_stage_idx, topk * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    BO_CUDA(h, cudaStreamSynchronize(st));
    return 0;
}

int bo_sobol_points(bo_handle* h, const bo_sobol* sobol_host, const int64_t* idx_dev, int64_t N, double* out_dev,
                    void* stream) {
    if (!h) return BO_E_INVALID;
    return sobol_points_impl(h, sobol_host, idx_dev, N, out_dev, (cudaStream_t)stream);
}

int bo_refine(bo_handle* h, int32_t acq_kind, double best_f, double beta, double min_variance,
              const double* starts_dev, int32_t k, int32_t iters, double* x_dev, double* val_dev, void* stream) {
    if (!h) return BO_E_INVALID;
    return refine_impl(h, acq_kind, best_f, beta, min_variance, starts_dev, k, iters, x_dev, val_dev, (cudaStream_t)stream);
}

int bo_acq_grad(bo_handle* h, int32_t acq_kind, double best_f, double beta, double min_variance,
                const double* Xq_dev, int32_t k, double* val_dev, double* grad_dev, void* stream) {
    if (!h) return BO_E_INVALID;
    return acq_grad_impl(h, acq_kind, best_f, beta, min_variance, Xq_dev, k, val_dev, grad_dev, (cudaStream_t)stream);
}

int bo_append(bo_handle* h, const double* x_dev, double y, int32_t use_believer, void* stream) {
    if (!h) return BO_E_INVALID;
    return append_impl(h, x_dev, y, use_believer, (cudaStream_t)stream);
}

int bo_lml_grad_batched(bo_handle* h, const double* X_dev, const double* y_dev, int32_t n, int32_t d,
                        int32_t kernel_kind, double mean, const double* theta_host, int32_t R, double* lml_host,
                        double* grad_host, int32_t* status_host, void* stream) {
    if (!h) return BO_E_INVALID;
    return lml_impl(h, X_dev, y_dev, n, d, kernel_kind, mean, theta_host, R, lml_host, grad_host, status_host,
                    (cudaStream_t)stream);
}

int bo_fp64_peak(bo_handle* h, int32_t use_dmma, double seconds, double* tflops_host) {
    if (!h || !tflops_host) return BO_E_INVALID;
    return fp64_peak_impl(h, use_dmma, seconds, tflops_host);
}

int bo_fps(bo_handle* h, const double* X_dev, int64_t N, int32_t d, int32_t m, int64_t start, int64_t* idx_dev, void* stream) {
    if (!h) return BO_E_INVALID;
    return fps_impl(h, X_dev, N, d, m, start, idx_dev, (cudaStream_t)stream);
}

int bo_gemm_probe(bo_handle* h, int32_t m, int32_t n, int32_t k, int32_t cfg, int32_t reps, double* tflops_host) {
    if (!h || !tflops_host || reps < 1) return BO_E_INVALID;
    return gemm_probe_impl(h, m, n, k, cfg, reps, tflops_host);
}

int bo_set_sweep_mode(bo_handle* h, int32_t mode) {
    if (!h) return BO_E_INVALID;
    if (mode < BO_SWEEP_AUTO || mode > BO_SWEEP_I8X8) return fail(h, BO_E_INVALID, "bo_set_sweep_mode: unknown mode");
    h->sweep_mode = mode;
    return 0;
}

int bo_resolve_sweep_mode(const bo_handle* h, int64_t pool_total) {
    if (!h) return BO_E_INVALID;
    if (!h->fitted) return BO_E_NOTFIT;
    return resolve_sweep_mode(h, h->sweep_mode, pool_total);
}

int bo_set_linear_variance_ard(bo_handle* h, const double* variances_host, int32_t d) {
    if (!h) return BO_E_INVALID;
    h->lin_v_ard.clear();
    if (!variances_host || d == 0) return 0;
    if (d < 0 || d > BO_MAX_DIM) return fail(h, BO_E_CAPACITY, "bo_set_linear_variance_ard: d exceeds BO_MAX_DIM");
    for (int k = 0; k < d; ++k)
        if (!(variances_host[k] >= 0.0)) return fail(h, BO_E_INVALID, "bo_set_linear_variance_ard: variances must be >= 0");
    h->lin_v_ard.assign(variances_host, variances_host + d);
    return 0;
}

int bo_last_sweep_path(const bo_handle* h) { return (h && h->sweep_timed) ? h->sweep_path : -1; }
int64_t bo_last_sweep_flagged(const bo_handle* h) { return (h && h->sweep_timed) ? h->sweep_flagged : 0; }

int bo_i8_peak(bo_handle* h, double seconds, double* tops_host) {
    if (!h || !tops_host) return BO_E_INVALID;
    return i8_peak_impl(h, seconds, tops_host);
}

double bo_last_sweep_ms(bo_handle* h) {
    if (!h || !h->sweep_timed) return -1.0;
    cudaSetDevice(h->device);
    if (cudaEventSynchronize(h->ev1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    return (double)ms;
}

}  // extern "C"
