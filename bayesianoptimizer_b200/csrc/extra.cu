// K5-K7 (refinement, Kriging-believer append, batched LML) and the FP64 peak probe.
#include "gemm.cuh"

namespace bo {

// ---- FP64 peak probe: register-resident DMMA.8x8x4 / DFMA loops (roofline denominator) ----------
__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, int iters, double a, double b) {
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

int fp64_peak_impl(bo_handle* h, int use_dmma, double seconds, double* tflops) {
    BO_CUDA(h, cudaSetDevice(h->device));
    const int blocks = h->sm_count * 8, threads = 256, iters = 20000;
    const double warps = (double)blocks * threads / 32;
    const double flop_per_launch = use_dmma ? warps * 8 * 2.0 * 256 * iters : warps * 32 * 8 * 2.0 * iters;
    cudaEvent_t e0, e1;
    BO_CUDA(h, cudaEventCreate(&e0));
    BO_CUDA(h, cudaEventCreate(&e1));
    auto launch = [&]() {
        if (use_dmma) peak_dmma_kernel<<<blocks, threads>>>(h->vec1 ? h->vec1 : (double*)h->info_dev, iters, 1.0000001, 1e-9);
        else peak_dfma_kernel<<<blocks, threads>>>(h->vec1 ? h->vec1 : (double*)h->info_dev, iters, 1.0000001, 1e-9);
        h->launches++;
    };
    launch();
    BO_CUDA(h, cudaDeviceSynchronize());
    // one launch is ~10 ms (DMMA) / ~3 ms (DFMA); run enough launches to cover `seconds`
    BO_CUDA(h, cudaEventRecord(e0));
    launch();
    BO_CUDA(h, cudaEventRecord(e1));
    BO_CUDA(h, cudaEventSynchronize(e1));
    float ms1 = 0.f;
    BO_CUDA(h, cudaEventElapsedTime(&ms1, e0, e1));
    int reps = (int)(seconds * 1e3 / (ms1 > 0.01f ? ms1 : 0.01f));
    if (reps < 1) reps = 1;
    if (reps > 2000) reps = 2000;
    BO_CUDA(h, cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    BO_CUDA(h, cudaEventRecord(e1));
    BO_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0.f;
    BO_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    BO_CUDA(h, cudaGetLastError());
    *tflops = flop_per_launch * reps / (ms * 1e-3) * 1e-12;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}


#define BO_DISPATCH_DP(dp, fn, ...)                                   \
    ((dp) == 2 ? fn<2>(__VA_ARGS__) : (dp) == 4 ? fn<4>(__VA_ARGS__) :  \
     (dp) == 6 ? fn<6>(__VA_ARGS__) : (dp) == 8 ? fn<8>(__VA_ARGS__) :  \
     (dp) == 12 ? fn<12>(__VA_ARGS__) : fn<16>(__VA_ARGS__))

// =================================================================================================
// K6: acquisition value + analytic gradient at k query points, and batched refinement of starts.
// Replaces the autograd backward of acq(X) and gen_candidates_scipy inside optimize_acqf
// (optimization/Bayesian.py:105-112).
// =================================================================================================

// kv[q][j] = k(x_q, X_j), gv[q][j] = "g" with dk/dx = g * (x~ - X~_j) * inv_ls   (0 for j >= n)
template <int DP>
__global__ void __launch_bounds__(256) kq_build_kernel(const double* __restrict__ Xs, int n, int np, Hyper hyp,
                                                       const double* __restrict__ Xq, int d,
                                                       double* __restrict__ kv, double* __restrict__ gv) {
    const int j = blockIdx.x * 256 + threadIdx.x, qi = blockIdx.y;
    if (j >= np) return;
    double kval = 0.0, g = 0.0;
    if (j < n) {
        double sq = 0.0, lin = 0.0;
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            const double xq = (k < d) ? Xq[(size_t)qi * d + k] * hyp.inv_ls[k] : 0.0;
            const double xj = Xs[(size_t)j * BO_MAX_DIM + k];
            const double df = xq - xj;
            sq = fma(df, df, sq);
            lin = fma(hyp.lin_w[k] * xq, xj, lin);
        }
        if (hyp.kind == BO_KERNEL_MATERN52 || hyp.kind == BO_KERNEL_LINEAR_MATERN52) {
            const double s5 = 2.23606797749978969640917366873128;
            const double r = sqrt(sq), e = exp(-s5 * r);
            kval = hyp.outputscale * fma(sq, 5.0 / 3.0, fma(s5, r, 1.0)) * e;
            g = -hyp.outputscale * (5.0 / 3.0) * fma(s5, r, 1.0) * e;
            // linear + Matern kind: k = s2 (v <x, x'> + matern); g stays the Matern factor, the linear term's own
            // derivative s2 v x'_k is added where the gradient is assembled (acq_grad_finalize_kernel)
            if (hyp.kind == BO_KERNEL_LINEAR_MATERN52) kval = fma(hyp.outputscale, lin, kval);
        } else {
            kval = hyp.outputscale * exp(-0.5 * sq);
            g = -kval;
        }
    }
    kv[(size_t)qi * np + j] = kval;
    if (gv) gv[(size_t)qi * np + j] = g;
}

// V[q][i] = sum_{j<=i} Li[i][j] rhs[q][j] for QB right-hand sides at once (warp per row)
template <int QB>
__global__ void __launch_bounds__(256) trmm_lower_skinny_kernel(const double* __restrict__ Li, int ld, int np,
                                                                const double* __restrict__ rhs, int k,
                                                                double* __restrict__ out) {
    __shared__ double sh[QB][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * 8 + warp, q0 = blockIdx.y * QB;
    const int last = blockIdx.x * 8 + 7;
    double acc[QB];
#pragma unroll
    for (int q = 0; q < QB; ++q) acc[q] = 0.0;
    for (int j0 = 0; j0 <= last; j0 += 256) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < QB; ++q)
            sh[q][threadIdx.x] = (q0 + q < k && j0 + threadIdx.x < np) ? rhs[(size_t)(q0 + q) * np + j0 + threadIdx.x] : 0.0;
        __syncthreads();
        for (int jj = lane; jj < 256; jj += 32) {
            const int j = j0 + jj;
            if (j <= row) {
                const double a = Li[(size_t)row * ld + j];
#pragma unroll
                for (int q = 0; q < QB; ++q) acc[q] = fma(a, sh[q][jj], acc[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < QB; ++q) {
        double v = acc[q];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && q0 + q < k) out[(size_t)(q0 + q) * np + row] = v;
    }
}

// Wp[s][q][j] = sum over the s-th row segment of { i >= j } of Li[i][j] V[q][i] for QB right-hand sides
// (block per 32 columns x row segment x right-hand-side group; the finaliser adds the segments in order).
// The rows below a column block are split into `gridDim.y` contiguous segments so that the long first columns do
// not serialise on a single CTA (one CTA walking all np rows took 0.5 ms at n = 3000).
template <int QB>
__global__ void __launch_bounds__(256) trmm_lower_t_skinny_kernel(const double* __restrict__ Li, int ld, int np,
                                                                  const double* __restrict__ V, int k,
                                                                  double* __restrict__ out) {
    __shared__ double red[8][QB][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j0 = blockIdx.x * 32, j = j0 + tx, q0 = blockIdx.z * QB;
    const int rows = np - j0;
    const int chunk = ((rows + (int)gridDim.y - 1) / (int)gridDim.y + 7) & ~7;
    const int r0 = j0 + blockIdx.y * chunk, r1 = min(np, r0 + chunk);
    double acc[QB];
#pragma unroll
    for (int q = 0; q < QB; ++q) acc[q] = 0.0;
#pragma unroll 4
    for (int i = r0 + ty; i < r1; i += 8) {
        const double a = (i >= j) ? Li[(size_t)i * ld + j] : 0.0;
#pragma unroll
        for (int q = 0; q < QB; ++q)
            if (q0 + q < k) acc[q] = fma(a, V[(size_t)(q0 + q) * np + i], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < QB; ++q) red[ty][q][tx] = acc[q];
    __syncthreads();
    if (ty == 0) {
        double* o = out + (size_t)blockIdx.y * k * np;
#pragma unroll
        for (int q = 0; q < QB; ++q) {
            double t = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) t += red[r][q][tx];
            if (q0 + q < k) o[(size_t)(q0 + q) * np + j] = t;
        }
    }
}

__device__ __forceinline__ double block_sum_256(double v, double* red) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    return t;
}

__device__ __forceinline__ double log1mexp_e(double x) {
    return (x > -0.69314718055994530942) ? log(-expm1(x)) : log1p(-exp(x));
}

// per query: mu, var, grad mu, grad var -> acquisition value and gradient (block per query)
template <int DP>
__global__ void __launch_bounds__(256) acq_grad_finalize_kernel(const double* __restrict__ Xs, const double* __restrict__ alpha,
                                                                int n, int np, Hyper hyp, const double* __restrict__ Xq, int d,
                                                                const double* __restrict__ kv, const double* __restrict__ gv,
                                                                const double* __restrict__ V, const double* __restrict__ W, int wsplits, int k,
                                                                int acq, double best_f, double sqrt_beta, double min_var,
                                                                double* __restrict__ val, double* __restrict__ grad) {
    __shared__ double red[8];
    const int qi = blockIdx.x, tid = threadIdx.x;
    double xq[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) xq[k] = (k < d) ? Xq[(size_t)qi * d + k] * hyp.inv_ls[k] : 0.0;
    double mu = 0.0, ss = 0.0, gm[DP], gs[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) gm[k] = gs[k] = 0.0;
    for (int j = tid; j < np; j += 256) {
        const double v = V[(size_t)qi * np + j];
        ss = fma(v, v, ss);
        if (j < n) {
            double w = 0.0;
            for (int sp = 0; sp < wsplits; ++sp) w += W[((size_t)sp * k + qi) * np + j];      // row segments, in order
            const double a = alpha[j], g = gv[(size_t)qi * np + j];
            mu = fma(kv[(size_t)qi * np + j], a, mu);
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                const double xj = Xs[(size_t)j * BO_MAX_DIM + k];
                // dk_j/dx_k = g (x~_k - X~_jk) / l_k  [+ s2 v X_jk for the linear + Matern kind: lin_w = v l^2 on scaled inputs]
                double t = g * (xq[k] - xj) * hyp.inv_ls[k];
                if (hyp.kind == BO_KERNEL_LINEAR_MATERN52) t = fma(hyp.outputscale * hyp.lin_w[k] * xj, hyp.inv_ls[k], t);
                gm[k] = fma(a, t, gm[k]);
                gs[k] = fma(w, t, gs[k]);
            }
        }
    }
    mu = block_sum_256(mu, red);
    ss = block_sum_256(ss, red);
#pragma unroll
    for (int k = 0; k < DP; ++k) { gm[k] = block_sum_256(gm[k], red); gs[k] = block_sum_256(gs[k], red); }
    if (tid != 0) return;
    const double mean = hyp.mean + mu;
    double prior = hyp.outputscale;
    double dvar[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) dvar[k] = -2.0 * gs[k];
    if (hyp.kind == BO_KERNEL_LINEAR_MATERN52) {
        // prior variance s2 (v |x|^2 + 1) moves with x: d/dx_k = 2 s2 v x_k
        double nn = 0.0;
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            nn = fma(hyp.lin_w[k] * xq[k], xq[k], nn);
            dvar[k] = fma(2.0 * hyp.outputscale * hyp.lin_w[k] * xq[k], hyp.inv_ls[k], dvar[k]);
        }
        prior = hyp.outputscale * (nn + 1.0);
    }
    double var = prior - ss;
    if (var < min_var) {
        var = min_var;
#pragma unroll
        for (int k = 0; k < DP; ++k) dvar[k] = 0.0;
    }
    const double sigma = sqrt(var);
    double value, cm, cs;          // gradient = cm * dmu + cs * dsigma, dsigma = dvar / (2 sigma)
    const double inv_sqrt2 = 0.70710678118654752440, inv_sqrt_2pi = 0.39894228040143267794;
    if (acq == BO_ACQ_VAR) { value = var; cm = 0.0; cs = 2.0 * sigma; }
    else if (acq == BO_ACQ_MEAN) { value = mean; cm = 1.0; cs = 0.0; }
    else if (acq == BO_ACQ_UCB) { value = fma(sqrt_beta, sigma, mean); cm = 1.0; cs = sqrt_beta; }
    else {
        const double u = (mean - best_f) / sigma;
        const double phi = inv_sqrt_2pi * exp(-0.5 * u * u);
        const double Phi = 0.5 * erfc(-u * inv_sqrt2);
        if (acq == BO_ACQ_EI) {
            value = sigma * fma(u, Phi, phi); cm = Phi; cs = phi;
        } else {
            double lh, dlh;      // log h(u), d log h / du
            if (u > -1.0) {
                const double hh = fma(u, Phi, phi);
                lh = log(hh); dlh = Phi / hh;
            } else {
                const double ex = erfcx(-u * inv_sqrt2);
                const double t = 1.2533141373155002512 * ex;                 // sqrt(pi/2) erfcx(|u|/sqrt2) = Phi/phi
                const double w = log(ex * fabs(u)) + 0.22579135264472743236;
                lh = -0.5 * u * u - 0.91893853320467274178 + log1mexp_e(w);
                dlh = t / (-expm1(w));                                        // Phi / (phi + u Phi)
            }
            value = log(sigma) + lh;
            // d/dx = dsigma/sigma + dlh * (dmu - u dsigma)/sigma
            cm = dlh / sigma; cs = (1.0 - dlh * u) / sigma;
        }
    }
    val[qi] = value;
#pragma unroll
    for (int k = 0; k < DP; ++k)
        if (k < d) grad[(size_t)qi * d + k] = cm * gm[k] + cs * dvar[k] / (2.0 * sigma);
}

static int ensure_qbuf(bo_handle* h, size_t elems) {
    if (elems <= h->qbuf_elems) return 0;
    if (h->qbuf) cudaFree(h->qbuf);
    h->qbuf = nullptr; h->qbuf_elems = 0;
    BO_CUDA(h, cudaMalloc(&h->qbuf, elems * sizeof(double)));
    h->qbuf_elems = elems;
    return 0;
}

// right-hand sides per pass over L^-1: 8.  (16 would cover the default 10 restarts in one pass, but the kernels are bound by
// the per-element FMA/load chains, not by the L^-1 stream: measured 75 vs 31 us for the transposed product at n = 3000.)
static int acq_qb(int k) { (void)k; return 8; }
// row segments of the transposed skinny TRMM: enough CTAs for ~8 waves, at most 16 (workspace = (3 + splits) k np doubles)
static int acq_wsplits(const bo_handle* h, int k) {
    const int qb = acq_qb(k);
    const int base = (h->np / 32) * ((k + qb - 1) / qb);
    int s = (8 * h->sm_count + base - 1) / base;
    return s < 1 ? 1 : (s > 16 ? 16 : s);
}
static size_t acq_ws_elems(const bo_handle* h, int k) { return (size_t)(3 + acq_wsplits(h, k)) * k * h->np; }

template <int DP>
static int eval_acq_grad(bo_handle* h, int acq, double best_f, double beta, double min_var, const double* Xq, int k,
                         double* val, double* grad, double* ws, cudaStream_t st) {
    const int np = h->np, ld = h->cap_np;
    double* kv = ws; double* gv = kv + (size_t)k * np; double* V = gv + (size_t)k * np; double* W = V + (size_t)k * np;
    kq_build_kernel<DP><<<dim3(np / 256 + (np % 256 != 0), k), 256, 0, st>>>(h->Xs, h->n, np, h->hyp, Xq, h->d, kv, gv);
    BO_LAUNCH_CHECK(h);
    const int qb = acq_qb(k), qg = (k + qb - 1) / qb;
    const int ws_splits = acq_wsplits(h, k);
    if (qb == 16) {
        trmm_lower_skinny_kernel<16><<<dim3(np / 8, qg), 256, 0, st>>>(h->Li, ld, np, kv, k, V);
        BO_LAUNCH_CHECK(h);
        trmm_lower_t_skinny_kernel<16><<<dim3(np / 32, ws_splits, qg), 256, 0, st>>>(h->Li, ld, np, V, k, W);
    } else {
        trmm_lower_skinny_kernel<8><<<dim3(np / 8, qg), 256, 0, st>>>(h->Li, ld, np, kv, k, V);
        BO_LAUNCH_CHECK(h);
        trmm_lower_t_skinny_kernel<8><<<dim3(np / 32, ws_splits, qg), 256, 0, st>>>(h->Li, ld, np, V, k, W);
    }
    BO_LAUNCH_CHECK(h);
    acq_grad_finalize_kernel<DP><<<k, 256, 0, st>>>(h->Xs, h->alpha, h->n, np, h->hyp, Xq, h->d, kv, gv, V, W, ws_splits, k, acq, best_f,
                                                   sqrt(beta), min_var, val, grad);
    BO_LAUNCH_CHECK(h);
    return 0;
}

static int check_query_args(bo_handle* h, int acq_kind, double beta, const void* a, const void* b, const void* c, int k) {
    if (!h->fitted) return fail(h, BO_E_NOTFIT, "acquisition gradient / refinement before a successful bo_fit");
    if (h->svgp) return fail(h, BO_E_INVALID, "acquisition gradients / refinement are not defined on an SVGP predictive state");
    if (acq_kind < BO_ACQ_EI || acq_kind > BO_ACQ_MEAN) return fail(h, BO_E_INVALID, "unknown acquisition kind");
    if (!a || !b || !c || k < 1) return fail(h, BO_E_INVALID, "bad query arguments");
    if (k > 4096) return fail(h, BO_E_CAPACITY, "at most 4096 query points per call");
    if (!(beta >= 0.0)) return fail(h, BO_E_INVALID, "beta must be >= 0");
    return 0;
}

int acq_grad_impl(bo_handle* h, int acq_kind, double best_f, double beta, double min_var, const double* Xq_dev, int k,
                  double* val_dev, double* grad_dev, cudaStream_t st) {
    int rc = check_query_args(h, acq_kind, beta, Xq_dev, val_dev, grad_dev, k);
    if (rc) return rc;
    BO_CUDA(h, cudaSetDevice(h->device));
    if ((rc = ensure_qbuf(h, acq_ws_elems(h, k) + 64))) return rc;
    return BO_DISPATCH_DP(h->dp, eval_acq_grad, h, acq_kind, best_f, beta, min_var, Xq_dev, k, val_dev, grad_dev, h->qbuf, st);
}

// ---- refinement state machine (per start; projected ascent with Barzilai-Borwein steps) ---------
__global__ void refine_init_kernel(int k, int d, const double* __restrict__ g, double* __restrict__ step) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= k) return;
    double gmax = 0.0;
    for (int c = 0; c < d; ++c) gmax = fmax(gmax, fabs(g[(size_t)s * d + c]));
    step[s] = (gmax > 0.0 && isfinite(gmax)) ? 0.05 / gmax : 0.0;
}
__global__ void refine_propose_kernel(int k, int d, const double* __restrict__ x, const double* __restrict__ g,
                                      const double* __restrict__ step, double* __restrict__ xn) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= k * d) return;
    const int s = e / d;
    double gv = g[e];
    if (!isfinite(gv)) gv = 0.0;
    xn[e] = fmin(1.0, fmax(0.0, fma(step[s], gv, x[e])));
}
__global__ void refine_update_kernel(int k, int d, double* __restrict__ x, double* __restrict__ f, double* __restrict__ g,
                                     const double* __restrict__ xn, const double* __restrict__ fn,
                                     const double* __restrict__ gn, double* __restrict__ step) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= k) return;
    const double f_new = fn[s];
    if (f_new > f[s] && isfinite(f_new)) {
        double ss = 0.0, sy = 0.0;
        for (int c = 0; c < d; ++c) {
            const double sd = xn[(size_t)s * d + c] - x[(size_t)s * d + c];
            const double yd = gn[(size_t)s * d + c] - g[(size_t)s * d + c];
            ss = fma(sd, sd, ss); sy = fma(sd, yd, sy);
            x[(size_t)s * d + c] = xn[(size_t)s * d + c];
            g[(size_t)s * d + c] = gn[(size_t)s * d + c];
        }
        f[s] = f_new;
        double st = step[s];
        st = (sy < 0.0) ? -ss / sy : st * 2.0;           // BB1 step on negative curvature, else expand
        step[s] = fmin(fmax(st, 1e-12), 1e6);
    } else {
        step[s] *= 0.25;                                   // reject: shrink and retry from the same point
    }
}

// number of starts that are NOT yet converged: projected gradient (components pushing out of the box dropped) below
// gtol * max(1, |f|), or a step so small that the point can no longer move (pgtol-style test of L-BFGS-B)
__global__ void refine_active_kernel(int k, int d, const double* __restrict__ x, const double* __restrict__ f,
                                     const double* __restrict__ g, const double* __restrict__ step, double gtol,
                                     int* __restrict__ active) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= k) return;
    double pg = 0.0;
    for (int c = 0; c < d; ++c) {
        const double xv = x[(size_t)s * d + c], gv = g[(size_t)s * d + c];
        const bool blocked = (xv <= 0.0 && gv < 0.0) || (xv >= 1.0 && gv > 0.0);
        if (!blocked && isfinite(gv)) pg = fmax(pg, fabs(gv));
    }
    const double fs = isfinite(f[s]) ? fabs(f[s]) : 0.0;
    if (pg > gtol * fmax(1.0, fs) && step[s] * pg > 1e-13) atomicAdd(active, 1);
}

int refine_impl(bo_handle* h, int acq_kind, double best_f, double beta, double min_var, const double* starts_dev, int k,
                int iters, double* x_dev, double* val_dev, cudaStream_t st) {
    int rc = check_query_args(h, acq_kind, beta, starts_dev, x_dev, val_dev, k);
    if (rc) return rc;
    if (iters < 0) return fail(h, BO_E_INVALID, "bo_refine: iters must be >= 0");
    BO_CUDA(h, cudaSetDevice(h->device));
    const int d = h->d;
    const size_t ws_elems = acq_ws_elems(h, k);
    // layout: [acq-grad scratch | g | xn | fn | gn | step]
    if ((rc = ensure_qbuf(h, ws_elems + (size_t)k * (3 * d + 2) + 64))) return rc;
    double* ws = h->qbuf;
    double* g = ws + ws_elems; double* xn = g + (size_t)k * d; double* gn = xn + (size_t)k * d;
    double* fn = gn + (size_t)k * d; double* step = fn + k;
    BO_CUDA(h, cudaMemcpyAsync(x_dev, starts_dev, (size_t)k * d * 8, cudaMemcpyDeviceToDevice, st));
    if ((rc = BO_DISPATCH_DP(h->dp, eval_acq_grad, h, acq_kind, best_f, beta, min_var, x_dev, k, val_dev, g, ws, st))) return rc;
    refine_init_kernel<<<(k + 127) / 128, 128, 0, st>>>(k, d, g, step);
    BO_LAUNCH_CHECK(h);
    constexpr int CHECK_EVERY = 16;          // convergence poll: one 4-byte read-back per 16 iterations
    for (int it = 0; it < iters; ++it) {
        if (it > 0 && it % CHECK_EVERY == 0 && it + CHECK_EVERY <= iters) {
            BO_CUDA(h, cudaMemsetAsync(h->info_dev, 0, sizeof(int), st));
            refine_active_kernel<<<(k + 127) / 128, 128, 0, st>>>(k, d, x_dev, val_dev, g, step, 1e-6, h->info_dev);
            BO_LAUNCH_CHECK(h);
            BO_CUDA(h, cudaMemcpyAsync(h->info_host, h->info_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
            BO_CUDA(h, cudaStreamSynchronize(st));
            if (*h->info_host == 0) break;     // every start has converged (maxiter is an upper bound, Bayesian.py:111)
        }
        refine_propose_kernel<<<(k * d + 127) / 128, 128, 0, st>>>(k, d, x_dev, g, step, xn);
        BO_LAUNCH_CHECK(h);
        if ((rc = BO_DISPATCH_DP(h->dp, eval_acq_grad, h, acq_kind, best_f, beta, min_var, xn, k, fn, gn, ws, st))) return rc;
        refine_update_kernel<<<(k + 127) / 128, 128, 0, st>>>(k, d, x_dev, val_dev, g, xn, fn, gn, step);
        BO_LAUNCH_CHECK(h);
    }
    return 0;
}

// =================================================================================================
// K5: row append (Kriging believer / new observation) by bordering L and L^-1
// =================================================================================================
__global__ void pad_identity_kernel(double* __restrict__ Lm, double* __restrict__ Li, int ld, int row0, int rows,
                                    double* __restrict__ Xs, double* __restrict__ Xraw, double* __restrict__ yv,
                                    double* __restrict__ alpha) {
    // rows [row0, row0+rows): zero + unit diagonal; auxiliary vectors zero
    const int r = row0 + blockIdx.x;
    for (int c = threadIdx.x; c < row0 + rows; c += blockDim.x) {
        const double v = (c == r) ? 1.0 : 0.0;
        Lm[(size_t)r * ld + c] = v;
        Li[(size_t)r * ld + c] = v;
    }
    if (threadIdx.x < BO_MAX_DIM) { Xs[(size_t)r * BO_MAX_DIM + threadIdx.x] = 0.0; Xraw[(size_t)r * BO_MAX_DIM + threadIdx.x] = 0.0; }
    if (threadIdx.x == 0) { yv[r] = 0.0; alpha[r] = 0.0; }
}

__global__ void __launch_bounds__(1024) append_finalize_kernel(int n, int np, int ld, Hyper hyp, const double* __restrict__ x, int d,
                                                               double y, int believer, const double* __restrict__ kv,
                                                               const double* __restrict__ l, const double* __restrict__ upart, int splits,
                                                               double* __restrict__ Lm, double* __restrict__ Li,
                                                               double* __restrict__ alpha, double* __restrict__ Xs,
                                                               double* __restrict__ Xraw, double* __restrict__ yv,
                                                               double* __restrict__ Lp, int* __restrict__ info) {
    __shared__ double red[2][32];
    __shared__ double bc[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double ss = 0.0, ka = 0.0;
    for (int j = tid; j < n; j += 1024) { ss = fma(l[j], l[j], ss); ka = fma(kv[j], alpha[j], ka); }
#pragma unroll
    for (int o = 16; o; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); ka += __shfl_xor_sync(0xffffffffu, ka, o); }
    if (lane == 0) { red[0][warp] = ss; red[1][warp] = ka; }
    __syncthreads();
    if (warp == 0) {
        ss = red[0][lane]; ka = red[1][lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); ka += __shfl_xor_sync(0xffffffffu, ka, o); }
        if (lane == 0) { bc[0] = ss; bc[1] = ka; }
    }
    __syncthreads();
    ss = bc[0]; ka = bc[1];
    double prior = hyp.outputscale;
    if (hyp.kind == BO_KERNEL_LINEAR_MATERN52) {
        double nn = 0.0;
        for (int k = 0; k < d; ++k) { const double xs = x[k] * hyp.inv_ls[k]; nn = fma(hyp.lin_w[k] * xs, xs, nn); }   // sum_k v_k x_k^2
        prior = hyp.outputscale * (nn + 1.0);
    }
    const double lam2 = prior + hyp.noise + hyp.jitter - ss;
    if (!(lam2 > 0.0)) { if (tid == 0) *info = n + 1; return; }
    const double lam = sqrt(lam2), ilam = 1.0 / lam;
    const double ynew = believer ? hyp.mean + ka : y;
    const double znew = (ynew - hyp.mean - ka) * ilam;
    // u = L^-T l arrives as row-split partial sums (added in ascending split order, as trmv_reduce_kernel does); the new row of
    // L^-1 also goes straight into its place in the packed tiles of the FP64 sweep (pack_linv_kernel's layout) -- no separate
    // reduction and block-row repack launches on a chain whose cost is launch latency
    constexpr int KCH = SW_BM / SW_BK;
    const int ib = n / SW_BM, r = n % SW_BM;
    double* prow = Lp + (size_t)ib * (ib + 1) / 2 * KCH * SW_TILE;
    auto packed = [&](int j) -> double& {
        const int kc = j / SW_BK, k = j % SW_BK;
        return prow[(size_t)kc * SW_TILE + ((r >> 3) * (SW_BK / 8) + (k >> 3)) * 64 + ((r & 7) * 4 + (k & 3)) * 2 + ((k & 7) >> 2)];
    };
    for (int j = tid; j < n; j += 1024) {
        double uj = 0.0;
        for (int sp = 0; sp < splits; ++sp) uj += upart[(size_t)sp * np + j];
        const double li = -uj * ilam;
        Lm[(size_t)n * ld + j] = l[j];
        Li[(size_t)n * ld + j] = li;
        packed(j) = li;
        alpha[j] = fma(li, znew, alpha[j]);
    }
    if (tid == 0) {
        Lm[(size_t)n * ld + n] = lam;
        Li[(size_t)n * ld + n] = ilam;
        packed(n) = ilam;
        alpha[n] = znew * ilam;
        yv[n] = ynew;
    }
    if (tid < BO_MAX_DIM) {
        const double xv = (tid < d) ? x[tid] : 0.0;
        Xraw[(size_t)n * BO_MAX_DIM + tid] = xv;
        Xs[(size_t)n * BO_MAX_DIM + tid] = xv * hyp.inv_ls[tid];
    }
}

// kernels defined in fit.cu, re-declared here through small host wrappers
int launch_trmv_lower(bo_handle* h, const double* v, double* z, cudaStream_t st);
int launch_trmv_lower_t(bo_handle* h, const double* z, double* out, int accumulate, cudaStream_t st);
int launch_trmv_lower_t_partial(bo_handle* h, const double* z, const double** part, int* splits, cudaStream_t st);

template <int DP>
static int launch_kq1(bo_handle* h, const double* x, double* kv, cudaStream_t st) {
    kq_build_kernel<DP><<<dim3(h->np / 256 + (h->np % 256 != 0), 1), 256, 0, st>>>(h->Xs, h->n, h->np, h->hyp, x, h->d, kv, nullptr);
    BO_LAUNCH_CHECK(h);
    return 0;
}

int append_impl(bo_handle* h, const double* x_dev, double y, int use_believer, cudaStream_t st) {
    if (!h->fitted) return fail(h, BO_E_NOTFIT, "bo_append before a successful bo_fit");
    if (h->svgp) return fail(h, BO_E_INVALID, "bo_append is not defined on an SVGP predictive state");
    if (!x_dev) return fail(h, BO_E_INVALID, "bo_append: null point");
    BO_CUDA(h, cudaSetDevice(h->device));
    int rc;
    const int np_before = h->np;
    const bool opened = h->n == h->np;
    if (opened) {
        // open a new padded block row: identity on the diagonal, zeros elsewhere
        const int np_new = h->np + PAD;
        if ((rc = ensure_capacity(h, np_new, st))) return rc;
        pad_identity_kernel<<<PAD, 256, 0, st>>>(h->Lm, h->Li, h->cap_np, h->np, PAD, h->Xs, h->Xraw, h->yv, h->alpha);
        BO_LAUNCH_CHECK(h);
        h->np = np_new;
        h->plan_np = -1;
    }
    if ((rc = ensure_qbuf(h, (size_t)3 * h->np + 64))) return rc;
    double* kv = h->qbuf; double* l = kv + h->np; double* u = l + h->np;
    if ((rc = BO_DISPATCH_DP(h->dp, launch_kq1, h, x_dev, kv, st))) return rc;
    if ((rc = launch_trmv_lower(h, kv, l, st))) return rc;
    const double* upart = nullptr; int splits = 0;
    if ((rc = launch_trmv_lower_t_partial(h, l, &upart, &splits, st))) return rc;
    BO_CUDA(h, cudaMemsetAsync(h->info_dev, 0, sizeof(int), st));
    // a block row opened by this call is packed once as a whole (identity rows); after that each append writes its own row
    if (opened && (rc = pack_row_block(h, h->n / SW_BM, st))) return rc;
    append_finalize_kernel<<<1, 1024, 0, st>>>(h->n, h->np, h->cap_np, h->hyp, x_dev, h->d, y, use_believer, kv, l, upart, splits, h->Lm,
                                              h->Li, h->alpha, h->Xs, h->Xraw, h->yv, h->Lp, h->info_dev);
    BO_LAUNCH_CHECK(h);
    BO_CUDA(h, cudaMemcpyAsync(h->info_host, h->info_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    BO_CUDA(h, cudaStreamSynchronize(st));
    if (*h->info_host != 0) {
        h->err = "bo_append: bordered matrix not positive definite (duplicate point with zero noise?)";
        // the padded row was left untouched (identity) by the finalize kernel; repacked tiles are unchanged
        if (h->np != np_before) { h->np = np_before; h->plan_np = -1; }      // give the freshly opened block row back
        return *h->info_host;
    }
    h->n += 1;
    h->factor_epoch++;
    return 0;
}


// =================================================================================================
// N3: greedy farthest-point sampling on the device (batch diversification / inducing-point selection,
// optimization/Bayesian7.py:82-107, optimization/Bayesian6.py:88-107): m - 1 sequential rounds of
// "distance to the newest pick -> running minimum -> arg-max (first index on ties)".  One persistent CTA;
// the reference runs the same loop as m pairs of torch.cdist / argmax launches on the CPU.
// =================================================================================================
__global__ void __launch_bounds__(1024) fps_kernel(const double* __restrict__ X, long long N, int d, int m, long long start,
                                                   double* __restrict__ dist, long long* __restrict__ idx_out) {
    __shared__ double xp[BO_MAX_DIM];
    __shared__ double sv[32];
    __shared__ long long si[32];
    __shared__ long long cur_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long cur = start;
    if (tid == 0) idx_out[0] = start;
    for (long long i = tid; i < N; i += 1024) dist[i] = INFINITY;
    for (int round = 1; round < m; ++round) {
        if (tid < d) xp[tid] = X[cur * d + tid];
        __syncthreads();
        double bv = -1.0; long long bi = 0x7fffffffffffffffLL;
        for (long long i = tid; i < N; i += 1024) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) { const double df = X[i * d + k] - xp[k]; s = fma(df, df, s); }
            const double dm = fmin(dist[i], s);
            dist[i] = dm;
            if (dm > bv) { bv = dm; bi = i; }                 // ascending i per thread: first index wins ties
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bv = sv[lane]; bi = si[lane];
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { cur_s = bi; idx_out[round] = bi; }
        }
        __syncthreads();
        cur = cur_s;
    }
}

int fps_impl(bo_handle* h, const double* X_dev, int64_t N, int d, int m, int64_t start, int64_t* idx_dev, cudaStream_t st) {
    if (!X_dev || !idx_dev || N < 1 || d < 1 || d > BO_MAX_DIM || m < 1 || start < 0 || start >= N)
        return fail(h, BO_E_INVALID, "bo_fps: bad argument");
    if (m > N) return fail(h, BO_E_INVALID, "bo_fps: m exceeds the number of points");
    BO_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_qbuf(h, (size_t)N + 64);
    if (rc) return rc;
    fps_kernel<<<1, 1024, 0, st>>>(X_dev, N, d, m, start, h->qbuf, (long long*)idx_dev);
    BO_LAUNCH_CHECK(h);
    return 0;
}


// =================================================================================================
// N4: m outputs sharing the fitted kernel matrix -- one Cholesky, m right-hand sides
// =================================================================================================
__global__ void multi_resid_kernel(const double* __restrict__ Y, int n, int np, int m, int t, double mean, double* __restrict__ r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) r[i] = (i < n) ? Y[(size_t)i * m + t] - mean : 0.0;
}

// mean[c][t] = means[t] + sum_j k(x*_c, X_j) A[t][j]: one warp per candidate, kernel evaluations shared by the m outputs
template <int DP, int MO>
__global__ void __launch_bounds__(256) multi_mean_kernel(const double* __restrict__ Xs, int n, int np, Hyper hyp,
                                                         const double* __restrict__ Xq, int d, long long N,
                                                         const double* __restrict__ A /*[m][np]*/, int m, int t0,
                                                         const double* __restrict__ means, double* __restrict__ out /*[N][m]*/) {
    const long long c = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= N) return;
    double xq[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) xq[k] = (k < d) ? Xq[(size_t)c * d + k] * hyp.inv_ls[k] : 0.0;
    double acc[MO];
#pragma unroll
    for (int t = 0; t < MO; ++t) acc[t] = 0.0;
    for (int j = lane; j < n; j += 32) {
        double sq = 0.0, lin = 0.0;
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            const double xj = Xs[(size_t)j * BO_MAX_DIM + k];
            const double df = xq[k] - xj;
            sq = fma(df, df, sq);
            lin = fma(hyp.lin_w[k] * xq[k], xj, lin);
        }
        double kv = kernel_value(hyp.kind, sq, hyp.outputscale);
        if (hyp.kind == BO_KERNEL_LINEAR_MATERN52) kv = fma(hyp.outputscale, lin, kv);
#pragma unroll
        for (int t = 0; t < MO; ++t)
            if (t0 + t < m) acc[t] = fma(kv, A[(size_t)(t0 + t) * np + j], acc[t]);
    }
#pragma unroll
    for (int t = 0; t < MO; ++t) {
        double v = acc[t];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && t0 + t < m) out[(size_t)c * m + t0 + t] = means[t0 + t] + v;
    }
}

int solve_alpha_rhs(bo_handle* h, const double* r, double* alpha_out, cudaStream_t st);   // fit.cu

template <int DP>
static int launch_multi_mean(bo_handle* h, const double* Xq, long long N, const double* A, int m, const double* means,
                             double* out, cudaStream_t st) {
    for (int t0 = 0; t0 < m; t0 += 8) {
        multi_mean_kernel<DP, 8><<<(unsigned)((N + 7) / 8), 256, 0, st>>>(h->Xs, h->n, h->np, h->hyp, Xq, h->d, N, A, m, t0, means, out);
        BO_LAUNCH_CHECK(h);
    }
    return 0;
}

int posterior_multi_impl(bo_handle* h, const double* Y_dev, int m, const double* means_host, const double* Xs_dev, int64_t N,
                         double min_var, double* mean_dev, double* var_dev, cudaStream_t st) {
    if (!h->fitted) return fail(h, BO_E_NOTFIT, "bo_posterior_multi before a successful bo_fit");
    if (h->svgp) return fail(h, BO_E_INVALID, "bo_posterior_multi is not defined on an SVGP predictive state");
    if (!Y_dev || m < 1 || m > 64 || N < 0 || (N > 0 && (!Xs_dev || !mean_dev))) return fail(h, BO_E_INVALID, "bo_posterior_multi: bad argument");
    BO_CUDA(h, cudaSetDevice(h->device));
    const int np = h->np;
    int rc = ensure_qbuf(h, (size_t)(m + 1) * np + 128);
    if (rc) return rc;
    double* A = h->qbuf;                         // [m][np] alphas
    double* r = A + (size_t)m * np;              // residual staging
    double* means_dev = r + np;                  // [m]
    std::vector<double> means(m, 0.0);
    if (means_host) for (int t = 0; t < m; ++t) means[t] = means_host[t];
    BO_CUDA(h, cudaMemcpyAsync(means_dev, means.data(), m * sizeof(double), cudaMemcpyHostToDevice, st));
    BO_CUDA(h, cudaStreamSynchronize(st));       // `means` is a pageable temporary
    for (int t = 0; t < m; ++t) {
        multi_resid_kernel<<<(np + 255) / 256, 256, 0, st>>>(Y_dev, h->n, np, m, t, means[t], r);
        BO_LAUNCH_CHECK(h);
        if ((rc = solve_alpha_rhs(h, r, A + (size_t)t * np, st))) return rc;
    }
    if (N > 0 && (rc = BO_DISPATCH_DP(h->dp, launch_multi_mean, h, Xs_dev, N, A, m, means_dev, mean_dev, st))) return rc;
    if (var_dev && N > 0)
        return sweep_impl(h, BO_ACQ_MEAN, 0.0, 0.0, min_var, Xs_dev, nullptr, 0, N, 0, nullptr, nullptr, nullptr, var_dev, nullptr, st,
                          h->sweep_mode == BO_SWEEP_AUTO ? BO_SWEEP_FP64 : -1);
    return 0;
}

}  // namespace bo
