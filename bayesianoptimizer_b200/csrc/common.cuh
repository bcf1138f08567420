// Shared declarations for the B200 (sm_100a) GP surrogate + acquisition library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <algorithm>
#include <new>
#include <string>

#include "../../include/bo_b200.h"

namespace bo {

// ---- geometry of the packed operands (see DESIGN.md "Data layout in HBM") -----------------
constexpr int NB        = 64;    // Cholesky / triangular-inverse leaf block size
constexpr int PAD       = 128;   // n is padded to a multiple of PAD (identity on the padded diagonal)
constexpr int SW_BM     = 128;   // sweep: rows of L^-1 per CTA tile
constexpr int SW_BN     = 128;   // sweep: candidates per CTA tile
constexpr int SW_BK     = 32;    // sweep: contraction (observation) chunk per pipeline stage
constexpr int SW_TILE   = SW_BM * SW_BK;          // doubles per packed A (or B) stage tile = 4096 (32 KB)
constexpr int SW_STAGES = 3;
constexpr int SW_CONSUMER_WARPS = 8;
constexpr int SW_THREADS = SW_CONSUMER_WARPS * 32;         // 2 warps per SM sub-partition: 255-register budget

struct Hyper {
    int    kind;
    int    d;          // true input dimension
    int    dp;         // padded dimension used by the templated kernels
    double inv_ls[BO_MAX_DIM];
    double lin_w[BO_MAX_DIM];   // linear + Matern kind: v * l_k^2, so that v <x, x'> = sum_k lin_w[k] x~_k x~'_k on the scaled inputs
    double lin_v;               // LinearKernel variance v (0 for the stationary kinds)
    double outputscale;
    double noise;
    double mean;
    double jitter;
};

// one problem of the grouped FP64 GEMM (gemm.cuh):  C = alpha * A * op(B) + beta * C, row-major
struct GemmProblem {
    const double* A; const double* B; double* C;
    int M, N, K;
    int lda, ldb, ldc;
    double alpha, beta;
    int transB;       // 0: B is [K][N] ; 1: B is [N][K]
    int mode;         // GEMM_* structure flags
    int tiles_n;      // N / BN
    int tile_begin;   // first flattened tile of this problem
    int tile_end;
};
struct GemmLaunch { int first, count, tiles, cfg; };   // cfg: 0 = 64x64 tiles, 1 = 128x128

}  // namespace bo

struct bo_handle {
    int device = 0;
    int sm_count = 0;
    std::string err;
    int64_t launches = 0;

    // fitted state
    bool fitted = false;
    int n = 0, np = 0, d = 0, dp = 0;       // np = n padded to bo::PAD
    int cap_np = 0, cap_d = 0;              // allocated capacity
    bo::Hyper hyp{};
    double* Xs = nullptr;        // [cap_np, BO_MAX_DIM] scaled inputs X / lengthscale (padded dims zero)
    double* Xraw = nullptr;      // [cap_np, BO_MAX_DIM] unscaled copy (refit after append, lml)
    double* yv = nullptr;        // [cap_np] targets
    double* alpha = nullptr;     // [cap_np]
    double* Lm = nullptr;        // [cap_np, cap_np] K then L (lower), row-major with ld = cap_np
    double* Li = nullptr;        // [cap_np, cap_np] L^-1 (lower), row-major with ld = cap_np
    double* Tw = nullptr;        // [cap_np * cap_np / 2] triangular-inverse workspace
    double* Lp = nullptr;        // packed L^-1 tiles in DMMA fragment order (sweep A operand)
    double* vec1 = nullptr;      // [cap_np] scratch vectors
    double* vec2 = nullptr;
    double* vec3 = nullptr;
    int*    info_dev = nullptr;  // pivot status
    int*    info_host = nullptr; // pinned
    // fit plan: every grouped-GEMM launch of the factorisation + inverse for the current np
    std::vector<bo::GemmProblem> plan_probs;
    std::vector<bo::GemmLaunch>  plan_launches;
    bo::GemmProblem* plan_dev = nullptr;
    size_t plan_dev_cap = 0;
    int plan_np = -1;

    // SVGP predictive mode (bo_svgp_load): the handle holds the inducing-point factorisation instead of an exact fit
    bool svgp = false;
    double sv_add = 0.0;         // K_uu jitter (gpytorch adds it to k** as well) + likelihood noise
    double* Lp2 = nullptr;       // packed J Ls^T J (second triangular factor of the predictive variance)
    size_t  Lp2_elems = 0;
    double* svB = nullptr;       // B = Ls^T L^-1 (np x np, dense, pitch np): the sliced sweep contracts the stack [L^-1; B] in one pass
    size_t  svB_elems = 0;
    double* panel2 = nullptr;    // [grid, np/SW_BK, SW_TILE] row-reversed interp-term panels
    size_t  panel2_bytes = 0;

    // sweep workspaces
    double* panel = nullptr;     // [grid, np/SW_BK, SW_TILE] K(X*,X) panels, one per resident CTA
    size_t  panel_bytes = 0;
    void*   split_ws = nullptr; size_t split_bytes = 0;   // small pools: per-(block, segment) partial sums + counters
    double* part_val = nullptr;  // [grid, BO_MAX_TOPK]
    int64_t* part_idx = nullptr;
    int     part_grid = 0;
    bo_sobol* sobol_dev = nullptr;
    double* cand_stage = nullptr; size_t cand_stage_bytes = 0;   // host-candidate staging
    double* out_stage_val = nullptr; int64_t* out_stage_idx = nullptr;  // device top-k staging for *_host
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool sweep_timed = false;

    // INT8-sliced sweep (sweep_i8.cuh): slices of L^-1 in stage-tile order, power-of-two row scales, per-CTA int8 panels
    int8_t* Lp8 = nullptr; size_t Lp8_bytes = 0;
    double* rowscale = nullptr; size_t rowscale_cap = 0;
    int8_t* panel8 = nullptr; size_t panel8_bytes = 0;
    int sweep_mode = BO_SWEEP_AUTO;   // bo_set_sweep_mode
    int sweep_path = 0;               // contraction of the last sweep: 0 = FP64 DMMA, 7 / 8 = INT8 slices
    // the int8 pack of L^-1 is cached: factor_epoch moves whenever L^-1 changes (fit, append, SVGP load)
    uint64_t factor_epoch = 1, Lp8_epoch = 0; int Lp8_S = 0;
    // per-dimension LinearKernel variances for the NEXT fit / SVGP load of the linear + Matern kind (bo_set_linear_variance_ard)
    std::vector<double> lin_v_ard;
    // accuracy guard of the sliced sweep: candidates whose variance is too small for the slicing error bound are listed
    // here by the sweep kernel and re-scored on the FP64 contraction (sweep_i8.cuh)
    double* guard_dev = nullptr;              // [2]: max_i rowscale_i^2 (i + 1) as a double and its bit pattern (atomicMax)
    long long* flag_idx = nullptr; size_t flag_cap = 0;
    int* flag_count_dev = nullptr; int* flag_count_host = nullptr;   // host copy is pinned
    long long sweep_flagged = 0;      // candidates of the last sweep that went through the FP64 re-score

    void* select_ws = nullptr; size_t select_bytes = 0;  // large top-K (select.cu)

    // K5-K7 workspaces
    double* qbuf = nullptr; size_t qbuf_elems = 0;      // refinement / acq-grad scratch
    void* lml_batch = nullptr;                          // bo::LmlBatch: slot workspaces of the batched LML restarts (lml.cuh)
};

namespace bo {

#define BO_CUDA(h, call)                                                                  \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            char buf__[512];                                                              \
            snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,   \
                     cudaGetErrorString(e__));                                            \
            (h)->err = buf__;                                                             \
            return (e__ == cudaErrorMemoryAllocation) ? BO_E_NOMEM : BO_E_CUDA;           \
        }                                                                                 \
    } while (0)

#define BO_LAUNCH_CHECK(h)                                                                \
    do {                                                                                  \
        (h)->launches++;                                                                  \
        BO_CUDA(h, cudaGetLastError());                                                   \
    } while (0)

inline int fail(bo_handle* h, int code, const char* msg) {
    if (h) h->err = msg;
    return code;
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// padded dimension served by the templated kernels
inline int pad_dim(int d) {
    if (d <= 2) return 2;
    if (d <= 4) return 4;
    if (d <= 6) return 6;
    if (d <= 8) return 8;
    if (d <= 12) return 12;
    return 16;
}

// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- branch-free FP64 elementary functions for the kernel evaluations --------------------------
// CUDA's sqrt()/exp() carry slow-path branches that split every kernel evaluation into several basic
// blocks, so independent evaluations cannot be interleaved and the panel build runs latency-bound.
// These versions are straight-line code (MUFU seed + Newton / Cody-Waite + Taylor), accurate to ~1-2 ulp.

// sqrt(x) for x >= 0 (returns ~1e-150 for x == 0, which is 0 for every use here)
__device__ __forceinline__ double sqrt_pos(double x) {
    const double xs = x + 1e-300;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xs));      // MUFU.RSQ64H, ~2^-22 relative
    const double h = 0.5 * xs;
    y = y * fma(-h, y * y, 1.5);
    y = y * fma(-h, y * y, 1.5);
    const double r = xs * y;
    return fma(0.5 * y, fma(-r, r, xs), r);                        // Heron correction
}

// exp(-t) for t >= 0; exact zero beyond t = 700 (the true value is below 1e-304)
__device__ __forceinline__ double exp_neg(double t) {
    const double MAGIC = 6755399441055744.0;                        // 1.5 * 2^52: round-to-nearest-integer trick
    const double z = fma(-t, 1.4426950408889634074, MAGIC);         // n = rint(-t * log2(e)) in the low word
    const int n = __double2loint(z);
    const double nf = z - MAGIC;
    double f = fma(nf, -6.93147180369123816490e-01, -t);            // f = -t - n ln2 (Cody-Waite, hi part)
    f = fma(nf, -1.90821492927058770002e-10, f);                    //                          (lo part)
    double p = 1.6059043836821613e-10;                              // Taylor of exp(f), |f| <= ln2/2, degree 13
    p = fma(p, f, 2.08767569878681e-09);
    p = fma(p, f, 2.505210838544172e-08);
    p = fma(p, f, 2.755731922398589e-07);
    p = fma(p, f, 2.7557319223985893e-06);
    p = fma(p, f, 2.48015873015873e-05);
    p = fma(p, f, 1.984126984126984e-04);
    p = fma(p, f, 1.388888888888889e-03);
    p = fma(p, f, 8.333333333333333e-03);
    p = fma(p, f, 4.1666666666666664e-02);
    p = fma(p, f, 1.6666666666666666e-01);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    const double r = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));   // * 2^n, n in [-1010, 0]
    return (t > 700.0) ? 0.0 : r;
}

// ---- leaner variants for the sliced sweep's panel build, where every FP64 instruction is paid for with tensor-pipe idle
// time (FP64 and kind::i8 MMAs throttle each other: DESIGN section 4).  Same accuracy class (~1.5 ulp), 9 FP64 instructions
// fewer per kernel evaluation: one Newton step less in the square root (rsqrt.approx 2^-22 -> 2^-43 -> Heron ~2^-86), exp
// through a 16-entry table of 2^(j/16) (shared memory, conflict-free) + a degree-7 polynomial on |f| <= ln2/32 (remainder
// 1e-18), the output scale folded into the Matern polynomial.
__device__ __forceinline__ double sqrt_pos_fast(double x) {
    const double xs = x + 1e-300;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xs));
    y = y * fma(-0.5 * xs, y * y, 1.5);
    const double r = xs * y;
    return fma(0.5 * y, fma(-r, r, xs), r);
}
// EXP2_16[j] = 2^(j/16), correctly rounded
#define BO_EXP2_16_TABLE {1.0, 1.0442737824274138, 1.0905077326652577, 1.1387886347566916, 1.189207115002721, 1.241857812073484, \
                          1.2968395546510096, 1.3542555469368927, 1.4142135623730951, 1.4768261459394993, 1.5422108254079407, \
                          1.6104903319492543, 1.681792830507429, 1.7562521603732995, 1.8340080864093424, 1.9152065613971474}
__device__ __forceinline__ double exp_neg_fast(double t, const double* __restrict__ tab /* shared: 2^(j/16) */) {
    const double MAGIC = 6755399441055744.0;                        // 1.5 * 2^52
    const double z = fma(-t, 23.083120654223414, MAGIC);            // n = rint(-t * 16 log2(e)) in the low word
    const int n = __double2loint(z);
    const double nf = z - MAGIC;
    double f = fma(nf, -4.33216987730702385306e-02, -t);            // f = -t - n ln2/16 (Cody-Waite: hi part has 32 significant bits)
    f = fma(nf, -1.19263433079411731251e-11, f);                    //                     (lo part)
    double p = 1.98412698412698413e-04;                             // Taylor of exp(f), |f| <= ln2/32, degree 7
    p = fma(p, f, 1.38888888888888894e-03);
    p = fma(p, f, 8.33333333333333322e-03);
    p = fma(p, f, 4.16666666666666644e-02);
    p = fma(p, f, 1.66666666666666657e-01);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    p *= tab[n & 15];                                               // 2^((n mod 16)/16)
    const double r = __hiloint2double(__double2hiint(p) + ((n >> 4) << 20), __double2loint(p));   // * 2^floor(n/16), in [-1010, 0]
    return (t > 700.0) ? 0.0 : r;
}
template <int KIND>
__device__ __forceinline__ double kernel_value_fast_t(double sq, double outputscale, const double* __restrict__ tab) {
    if (KIND == BO_KERNEL_MATERN52) {
        const double s5 = 2.23606797749978969640917366873128;
        const double r = sqrt_pos_fast(sq);
        const double p = fma(sq, outputscale * (5.0 / 3.0), fma(outputscale * s5, r, outputscale));     // s2 (1 + sqrt5 r + 5/3 r^2): loop-invariant constants
        return p * exp_neg_fast(s5 * r, tab);
    }
    return outputscale * exp_neg_fast(0.5 * sq, tab);
}

// Matern-5/2 / RBF value from the scaled squared distance (KIND known at compile time)
template <int KIND>
__device__ __forceinline__ double kernel_value_t(double sq, double outputscale) {
    if (KIND == BO_KERNEL_MATERN52) {
        const double s5 = 2.23606797749978969640917366873128;
        const double r = sqrt_pos(sq);
        const double p = fma(sq, 5.0 / 3.0, fma(s5, r, 1.0));
        return outputscale * p * exp_neg(s5 * r);
    }
    return outputscale * exp_neg(0.5 * sq);
}
__device__ __forceinline__ double kernel_value(int kind, double sq, double outputscale) {
    return kind == BO_KERNEL_RBF ? kernel_value_t<BO_KERNEL_RBF>(sq, outputscale)
                                 : kernel_value_t<BO_KERNEL_MATERN52>(sq, outputscale);      // also the Matern part of kind 2
}
// full kernel from the scaled squared distance and the weighted inner product lin = sum_k lin_w[k] x~_k x~'_k
template <int KIND>
__device__ __forceinline__ double kernel_pair_t(double sq, double lin, double outputscale) {
    if (KIND == BO_KERNEL_LINEAR_MATERN52) return fma(outputscale, lin, kernel_value_t<BO_KERNEL_MATERN52>(sq, outputscale));
    return kernel_value_t<KIND>(sq, outputscale);
}

// better-than order of the top-k: value desc, index asc
__device__ __forceinline__ bool tk_better(double va, long long ia, double vb, long long ib) {
    return va > vb || (va == vb && ia < ib);
}

}  // namespace bo

// internal cross-file API
namespace bo {
int fit_impl(bo_handle* h, const double* X_dev, const double* y_dev, int n, int d, int kind,
             const double* ls_host, double outputscale, double noise, double mean, double jitter,
             double linear_variance, cudaStream_t st);
int ensure_capacity(bo_handle* h, int np, cudaStream_t st);
int sweep_impl(bo_handle* h, int acq_kind, double best_f, double beta, double min_var,
               const double* cand_dev, const bo_sobol* sobol_host, int64_t first_index, int64_t N,
               int topk, double* vals_dev, int64_t* idx_dev, double* mean_dev, double* var_dev,
               double* acq_dev, cudaStream_t st, int mode_override = -1);     // mode_override: a BO_SWEEP_* mode, -1 = the handle's
int sobol_points_impl(bo_handle* h, const bo_sobol* sobol_host, const int64_t* idx_dev, int64_t N,
                      double* out_dev, cudaStream_t st);
int fp64_peak_impl(bo_handle* h, int use_dmma, double seconds, double* tflops);
int i8_peak_impl(bo_handle* h, double seconds, double* tops);
int resolve_sweep_mode(const bo_handle* h, int mode, long long pool);
int refit_factor(bo_handle* h, cudaStream_t st);
int svgp_load_impl(bo_handle* h, const double* Z_dev, int M, int d, int kind, const double* ls_host, double outputscale,
                   double linear_variance, double mean, double noise, double jitter, const double* var_mean_dev,
                   const double* var_chol_dev, cudaStream_t st);
int create_handle(bo_handle** out, int device);
void lml_release(bo_handle* h);
int pack_row_block(bo_handle* h, int ib, cudaStream_t st);   // K build + Cholesky + inverse + alpha from h->Xs/h->yv
}  // namespace bo
