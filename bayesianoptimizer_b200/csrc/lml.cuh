// K7: batched exact log marginal likelihood + gradient over R hyper-parameter restarts
// (the ExactMarginalLogLikelihood closure fit_gpytorch_mll evaluates, optimization/Bayesian.py:92-93).
//
// Included at the end of fit.cu (same translation unit: it launches the fit kernels).  Restarts are
// processed in lock-step groups of S slots: every stage of the factorisation is ONE launch covering all
// slots (Gram: blockIdx.z; leaf Cholesky: one block per slot; TRSM / SYRK / inverse levels / L^-T L^-1:
// one grouped-GEMM launch listing S problems), so the latency-bound panel chain is paid once per group
// instead of once per restart and the trailing updates fill the GPU.
//   lml  = -1/2 r^T alpha - sum log L_ii - n/2 log 2 pi
//   dlml/dtheta = 1/2 tr(W dK/dtheta),  W = alpha alpha^T - K^-1,  K^-1 = L^-T L^-1
#pragma once

struct LmlBatch {
    int S = 0, np = 0, n = 0, d = 0, dp = 0;
    double *Xraw = nullptr, *yv = nullptr;                 // shared by all slots
    double *Xs = nullptr, *Lm = nullptr, *Li = nullptr, *Tw = nullptr, *Kw = nullptr;   // per slot
    double *alpha = nullptr, *v1 = nullptr, *v2 = nullptr, *v3 = nullptr, *tpart = nullptr, *gpart = nullptr, *out = nullptr;
    Hyper* hyps = nullptr; int* info = nullptr;
    Hyper* hyps_host = nullptr; double* out_host = nullptr; int* info_host = nullptr;     // pinned
    std::vector<GemmProblem> probs; std::vector<GemmLaunch> launches; GemmProblem* plan_dev = nullptr;
    // slot groups run their (panel -> SYRK -> panel ...) chains on separate streams, so one group's latency-bound panel
    // kernels hide behind another group's trailing updates
    static constexpr int MAX_GROUPS = 8;
    cudaStream_t gstream[MAX_GROUPS] = {};
    cudaEvent_t fork_ev = nullptr, join_ev[MAX_GROUPS] = {};
};

static void lml_batch_free(LmlBatch* b) {
    if (!b) return;
    void* dev[] = {b->Xraw, b->yv, b->Xs, b->Lm, b->Li, b->Tw, b->Kw, b->alpha, b->v1, b->v2, b->v3, b->tpart, b->gpart,
                   b->out, b->hyps, b->info, b->plan_dev};
    for (void* p : dev) if (p) cudaFree(p);
    if (b->hyps_host) cudaFreeHost(b->hyps_host);
    if (b->out_host) cudaFreeHost(b->out_host);
    if (b->info_host) cudaFreeHost(b->info_host);
    for (int g = 0; g < LmlBatch::MAX_GROUPS; ++g) {
        if (b->gstream[g]) cudaStreamDestroy(b->gstream[g]);
        if (b->join_ev[g]) cudaEventDestroy(b->join_ev[g]);
    }
    if (b->fork_ev) cudaEventDestroy(b->fork_ev);
    cudaGetLastError();
    delete b;
}
void lml_release(bo_handle* h) { lml_batch_free(static_cast<LmlBatch*>(h->lml_batch)); h->lml_batch = nullptr; }

// per GRAM_ROWS x 32 tile of the lower triangle: partial sums of W_ij dK_ij/dlog l_k (k < DP), W_ij K_ij, and on the
// diagonal W_ii, log L_ii, r_i alpha_i.  blockIdx.z = slot.  Lane = column j, warp = 16 consecutive rows processed four
// at a time with branch-free code (masked pairs carry w = 0), kernel kind resolved outside the loops, the tile's rows of
// X~ / alpha staged in shared memory, one barrier for the whole reduction.
template <int DP, int KIND>
__device__ __forceinline__ void lml_pair_rows(const double (*xs_i)[DP], const double* al_i, const double* xj, double aj, int i0,
                                              int j, int n, int ld, const Hyper& hyp, const double* __restrict__ Kinv, int warp,
                                              double* acc) {
    double xw[KIND == BO_KERNEL_LINEAR_MATERN52 ? DP : 1];
    if (KIND == BO_KERNEL_LINEAR_MATERN52) {
#pragma unroll
        for (int k = 0; k < DP; ++k) xw[k] = hyp.lin_w[k] * xj[k];
    }
    const double s5 = 2.23606797749978969640917366873128;
    for (int rb = 0; rb < 16; rb += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = warp * 16 + rb + u, i = i0 + r;
            const bool valid = i < n && j < i;                        // strict lower triangle inside the data (j < i < n)
            const double kin = Kinv[(size_t)i * ld + j];
            const double w = valid ? fma(al_i[r], aj, -kin) : 0.0;
            double sq = 0.0, lin = 0.0, df2[DP];
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                const double xi = xs_i[r][k];
                const double df = xi - xj[k];
                df2[k] = df * df;
                sq += df2[k];
                if (KIND == BO_KERNEL_LINEAR_MATERN52) lin = fma(xw[k], xi, lin);
            }
            double kval, G;
            if (KIND != BO_KERNEL_RBF) {
                const double rr = sqrt_pos(sq), e = exp_neg(s5 * rr);
                const double t = fma(s5, rr, 1.0);
                kval = hyp.outputscale * fma(sq, 5.0 / 3.0, t) * e;
                G = hyp.outputscale * (5.0 / 3.0) * t * e;
            } else {
                kval = hyp.outputscale * exp_neg(0.5 * sq);
                G = kval;
            }
            const double wg = w * G;
#pragma unroll
            for (int k = 0; k < DP; ++k) acc[k] = fma(wg, df2[k], acc[k]);
            if (KIND == BO_KERNEL_LINEAR_MATERN52) {
                const double klin = hyp.outputscale * lin;            // s2 v <x_i, x_j>
                acc[DP] = fma(w, kval + klin, acc[DP]);
                acc[DP + 4] = fma(w, klin, acc[DP + 4]);              // sum_{i>j} W_ij s2 v <x_i, x_j>
            } else {
                acc[DP] = fma(w, kval, acc[DP]);
            }
        }
    }
}

template <int DP>
__global__ void __launch_bounds__(256, 2) lml_grad_tile_kernel(const double* __restrict__ Xs_all, const double* __restrict__ alpha_all,
                                                            const double* __restrict__ Kinv_all, const double* __restrict__ Lm_all,
                                                            const double* __restrict__ yv, int ld, int n, int np,
                                                            const Hyper* __restrict__ hyps, double* __restrict__ part_all) {
    __shared__ double xs_i[GRAM_ROWS][DP];
    __shared__ double al_i[GRAM_ROWS];
    __shared__ double red[8][DP + 6];
    const int bj = blockIdx.x, bi = blockIdx.y;
    if (bj * 32 > bi * GRAM_ROWS + GRAM_ROWS - 1) return;             // tile entirely above the diagonal
    const size_t s = blockIdx.z;
    const Hyper& hyp = hyps[s];
    const double* Xs = Xs_all + s * np * BO_MAX_DIM;
    const double* alpha = alpha_all + s * np;
    const double* Kinv = Kinv_all + s * np * ld;
    const double* Lm = Lm_all + s * np * ld;
    const int tile = bi * gridDim.x + bj;
    double* part = part_all + (s * gridDim.x * gridDim.y + tile) * (DP + 6);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = bi * GRAM_ROWS, j = bj * 32 + lane;
    for (int e = tid; e < GRAM_ROWS * DP; e += 256) xs_i[e / DP][e % DP] = Xs[(size_t)(i0 + e / DP) * BO_MAX_DIM + e % DP];
    if (tid < GRAM_ROWS) al_i[tid] = alpha[i0 + tid];
    double xj[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) xj[k] = Xs[(size_t)j * BO_MAX_DIM + k];           // padded rows of X~ are zero
    const double aj = alpha[j];
    __syncthreads();
    double acc[DP + 6];
#pragma unroll
    for (int k = 0; k < DP + 6; ++k) acc[k] = 0.0;
    if (hyp.kind == BO_KERNEL_MATERN52) lml_pair_rows<DP, BO_KERNEL_MATERN52>(xs_i, al_i, xj, aj, i0, j, n, ld, hyp, Kinv, warp, acc);
    else if (hyp.kind == BO_KERNEL_RBF) lml_pair_rows<DP, BO_KERNEL_RBF>(xs_i, al_i, xj, aj, i0, j, n, ld, hyp, Kinv, warp, acc);
    else lml_pair_rows<DP, BO_KERNEL_LINEAR_MATERN52>(xs_i, al_i, xj, aj, i0, j, n, ld, hyp, Kinv, warp, acc);
    // diagonal entries of this tile (columns that are also rows of the tile): one warp, lane = its own (j, j)
    if (warp == 0 && j >= i0 && j < i0 + GRAM_ROWS && j < n) {
        const double w = fma(aj, aj, -Kinv[(size_t)j * ld + j]);
        acc[DP + 1] += w;                                   // trace(W)
        acc[DP + 2] += log(Lm[(size_t)j * ld + j]);         // log det / 2
        acc[DP + 3] += (yv[j] - hyp.mean) * aj;             // quadratic form
        if (hyp.kind == BO_KERNEL_LINEAR_MATERN52) {
            double nn = 0.0;
#pragma unroll
            for (int k = 0; k < DP; ++k) nn = fma(hyp.lin_w[k] * xj[k], xj[k], nn);
            acc[DP + 5] = fma(w, hyp.outputscale * nn, acc[DP + 5]);            // sum_i W_ii s2 v |x_i|^2
        }
    }
#pragma unroll
    for (int k = 0; k < DP + 6; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (tid < DP + 6) {
        double t = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) t += red[w8][tid];
        part[tid] = t;
    }
}

// deterministic final reduction: one block per slot sums the tile partials column by column
template <int DP>
__global__ void __launch_bounds__(256) lml_reduce_kernel(const double* __restrict__ part_all, int tiles, int ntx, int n, int d,
                                                         const Hyper* __restrict__ hyps, double* __restrict__ out_all) {
    __shared__ double red[8];
    __shared__ double tot[DP + 6];
    const size_t s = blockIdx.x;
    const Hyper& hyp = hyps[s];
    const double* part = part_all + s * tiles * (DP + 6);
    double* out = out_all + s * (BO_MAX_DIM + 4);
    for (int k = 0; k < DP + 6; ++k) {
        double v = 0.0;
        for (int t = threadIdx.x; t < tiles; t += 256)
            if ((t % ntx) * 32 <= (t / ntx) * GRAM_ROWS + GRAM_ROWS - 1) v += part[(size_t)t * (DP + 6) + k];     // tiles that ran
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w8 = 0; w8 < 8; ++w8) t += red[w8];
            tot[k] = t;
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const double trW = tot[DP + 1], logdet_half = tot[DP + 2], quad = tot[DP + 3];
    out[0] = -0.5 * quad - logdet_half - 0.5 * n * 1.83787706640934548356;        // log(2 pi)
    for (int k = 0; k < d; ++k) out[1 + k] = tot[k];                               // pairs i>j count twice in the 1/2 sum
    out[1 + d] = tot[DP] + 0.5 * (hyp.outputscale * trW + tot[DP + 5]);            // d / d log outputscale
    out[2 + d] = 0.5 * hyp.noise * trW;                                            // d / d log noise
    out[3 + d] = tot[DP + 4] + 0.5 * tot[DP + 5];                                  // d / d log linear variance (kind 2)
}

// slots [s0, s0 + S) of the batch
template <int DP>
static int lml_launch_gram(bo_handle* h, LmlBatch* b, int s0, int S, cudaStream_t st) {
    const size_t c = b->np;
    gram_batched_kernel<DP><<<dim3(b->np / 32, b->np / GRAM_ROWS, S), dim3(32, 8), 0, st>>>(b->Xs + s0 * c * BO_MAX_DIM, b->n, b->np, b->np,
                                                                                    b->hyps + s0, b->Lm + s0 * c * c);
    BO_LAUNCH_CHECK(h);
    return 0;
}
template <int DP>
static int lml_launch_grad(bo_handle* h, LmlBatch* b, int s0, int S, cudaStream_t st) {
    const int nt = b->np / 32, nty = b->np / GRAM_ROWS;
    const size_t c = b->np;
    double* gpart = b->gpart + (size_t)s0 * nt * nty * (DP + 6);
    lml_grad_tile_kernel<DP><<<dim3(nt, nty, S), 256, 0, st>>>(b->Xs + s0 * c * BO_MAX_DIM, b->alpha + s0 * c, b->Kw + s0 * c * c,
                                                              b->Lm + s0 * c * c, b->yv, b->np, b->n, b->np, b->hyps + s0, gpart);
    BO_LAUNCH_CHECK(h);
    lml_reduce_kernel<DP><<<S, 256, 0, st>>>(gpart, nt * nty, nt, b->n, b->d, b->hyps + s0, b->out + (size_t)s0 * (BO_MAX_DIM + 4));
    BO_LAUNCH_CHECK(h);
    return 0;
}

// (re)build the slot workspace and the launch plan for (n, d, S)
static int lml_prepare(bo_handle* h, int n, int d, int S, cudaStream_t st) {
    LmlBatch* b = static_cast<LmlBatch*>(h->lml_batch);
    const int np = round_up(n, PAD), dp = pad_dim(d);
    // a workspace with MORE slots than this call needs is kept (the launches take the active prefix): the lock-step
    // optimiser's line search shrinks and grows the restart count from call to call, and re-allocating gigabytes each
    // time cost more than the factorisations (8 -> 37 ms per evaluation at n = 3000)
    if (b && b->np == np && b->S >= S && b->dp == dp) { b->n = n; b->d = d; return 0; }
    BO_CUDA(h, cudaStreamSynchronize(st));
    lml_batch_free(b);
    h->lml_batch = nullptr;
    b = new (std::nothrow) LmlBatch();
    if (!b) return fail(h, BO_E_NOMEM, "bo_lml_grad_batched: out of host memory");
    b->S = S; b->np = np; b->n = n; b->d = d; b->dp = dp;
    const size_t c = np, mat = c * c, nt = c / 32;
    struct { double** p; size_t elems; } reqs[] = {
        {&b->Xraw, c * BO_MAX_DIM}, {&b->yv, c}, {&b->Xs, S * c * BO_MAX_DIM}, {&b->Lm, S * mat}, {&b->Li, S * mat},
        {&b->Tw, S * (mat / 4 + 64)}, {&b->Kw, S * mat}, {&b->alpha, S * c}, {&b->v1, S * c}, {&b->v2, S * c}, {&b->v3, S * c},
        {&b->tpart, (size_t)S * TRMVT_SPLITS * c}, {&b->gpart, S * nt * nt * (dp + 6)}, {&b->out, (size_t)S * (BO_MAX_DIM + 4)}};
    cudaError_t e = cudaSuccess;
    for (auto& r : reqs) if (e == cudaSuccess) e = cudaMalloc(r.p, r.elems * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&b->hyps, S * sizeof(Hyper));
    if (e == cudaSuccess) e = cudaMalloc(&b->info, S * sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost(&b->hyps_host, S * sizeof(Hyper));
    if (e == cudaSuccess) e = cudaMallocHost(&b->out_host, (size_t)S * (BO_MAX_DIM + 4) * sizeof(double));
    if (e == cudaSuccess) e = cudaMallocHost(&b->info_host, S * sizeof(int));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->fork_ev, cudaEventDisableTiming);
    for (int g = 0; g < LmlBatch::MAX_GROUPS && e == cudaSuccess; ++g) {
        e = cudaStreamCreateWithFlags(&b->gstream[g], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->join_ev[g], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        lml_batch_free(b);
        cudaGetLastError();
        return fail(h, BO_E_NOMEM, "bo_lml_grad_batched: cannot allocate the restart slots");
    }
    // launch plan: per block column one SYRK launch (the panel kernel factors + solves), per inverse level two, then K^-1;
    // every launch lists the S slots' problems
    const int ld = np, nb = np / NB;
    auto plan_push = [&](const GemmBatch& g) {
        GemmLaunch L{(int)b->probs.size(), (int)g.probs.size(), g.tiles, g.bm == 128 ? 1 : 0};
        b->probs.insert(b->probs.end(), g.probs.begin(), g.probs.end());
        b->launches.push_back(L);
    };
    for (int kb = 0; kb + 1 < nb; ++kb) {
        const int m = np - (kb + 1) * NB;
        long t128 = (long)(m / 128) * (m / 128 + 1) / 2 * S;
        GemmBatch syrk(pick_tile(h->sm_count, {m}, t128));
        for (int s = 0; s < S; ++s) add_trailing_update(syrk, b->Lm + (size_t)s * mat, ld, np, kb, nb);
        plan_push(syrk);
    }
    struct Node { int lo, mid, hi, depth; };
    std::vector<Node> nodes, stack;
    int maxdepth = 0;
    if (nb > 1) stack.push_back({0, nb / 2, nb, 0});
    while (!stack.empty()) {
        Node nd = stack.back(); stack.pop_back();
        nodes.push_back(nd);
        if (nd.depth > maxdepth) maxdepth = nd.depth;
        if (nd.mid - nd.lo > 1) stack.push_back({nd.lo, nd.lo + (nd.mid - nd.lo) / 2, nd.mid, nd.depth + 1});
        if (nd.hi - nd.mid > 1) stack.push_back({nd.mid, nd.mid + (nd.hi - nd.mid) / 2, nd.hi, nd.depth + 1});
    }
    for (int depth = maxdepth; depth >= 0 && nb > 1; --depth) {
        bool all128 = true; long t128 = 0;
        for (const Node& nd : nodes) if (nd.depth == depth) {
            int p = (nd.mid - nd.lo) * NB, q = (nd.hi - nd.mid) * NB;
            if (p % 128 || q % 128 || (nd.lo * NB) % 128) all128 = false;
            t128 += (long)(p / 128) * (q / 128) * S;
        }
        const int tile = pick_tile(h->sm_count, {all128 ? 128 : 64}, t128);
        GemmBatch g1(tile), g2(tile);
        for (int s = 0; s < S; ++s) {
            double* Lm = b->Lm + (size_t)s * mat; double* Li = b->Li + (size_t)s * mat;
            size_t toff = (size_t)s * (mat / 4 + 64);
            for (const Node& nd : nodes) if (nd.depth == depth) {
                const int lo = nd.lo * NB, mid = nd.mid * NB, p = (nd.mid - nd.lo) * NB, q = (nd.hi - nd.mid) * NB;
                double* T = b->Tw + toff;
                g1.add(Lm + (size_t)mid * ld + lo, ld, Li + (size_t)lo * ld + lo, ld, T, p, q, p, p, 1.0, 0.0, 0, GEMM_B_LOWER_NN);
                g2.add(Li + (size_t)mid * ld + mid, ld, T, p, Li + (size_t)mid * ld + lo, ld, q, p, q, -1.0, 0.0, 0, GEMM_A_LOWER);
                toff += (size_t)q * p;
            }
        }
        plan_push(g1);
        plan_push(g2);
    }
    {
        GemmBatch kinv(pick_tile(h->sm_count, {np}, (long)(np / 128) * (np / 128 + 1) / 2 * S));
        for (int s = 0; s < S; ++s) {
            double* Li = b->Li + (size_t)s * mat;
            kinv.add(Li, ld, Li, ld, b->Kw + (size_t)s * mat, ld, np, np, np, 1.0, 0.0, 0, GEMM_TRANS_A | GEMM_LOWER_C | GEMM_K_FROM_MAX);
        }
        plan_push(kinv);
    }
    const size_t bytes = b->probs.size() * sizeof(GemmProblem);
    if (cudaMalloc(&b->plan_dev, bytes + 256) != cudaSuccess) { lml_batch_free(b); cudaGetLastError(); return fail(h, BO_E_NOMEM, "bo_lml_grad_batched: plan allocation failed"); }
    BO_CUDA(h, cudaMemcpy(b->plan_dev, b->probs.data(), bytes, cudaMemcpyHostToDevice));
    h->lml_batch = b;
    return 0;
}

static int lml_gemm(bo_handle* h, LmlBatch* b, int li, int s0, int S_active, cudaStream_t st) {
    // a launch lists the problems slot-major with identical tile counts per slot -> slots [s0, s0 + S_active) are a
    // contiguous sub-range of the list; the problems' tile_begin stay launch-absolute, hence the tile offset
    const GemmLaunch& L = b->launches[li];
    if (L.tiles == 0) return 0;
    const int per_slot = L.count / b->S, tiles_per_slot = L.tiles / b->S;
    const int count = per_slot * S_active, tiles = tiles_per_slot * S_active;
    const GemmProblem* pr = b->plan_dev + L.first + (size_t)per_slot * s0;
    if (L.cfg == 1) dgemm_grouped_kernel<128, 128><<<tiles, 256, GemmSmem<128, 128>::BYTES, st>>>(pr, count, tiles_per_slot * s0);
    else dgemm_grouped_kernel<64, 64, 16, 2, 4><<<tiles, 256, GemmSmem<64, 64, 16, 2>::BYTES, st>>>(pr, count, tiles_per_slot * s0);
    BO_LAUNCH_CHECK(h);
    return 0;
}

// the whole evaluation of slots [s0, s0 + Sa): Gram -> Cholesky -> inverse -> K^-1 -> alpha -> gradient partials
static int lml_run_slots(bo_handle* h, LmlBatch* b, int s0, int Sa, int n, int d, double mean, cudaStream_t st) {
    const int np = b->np, ld = np, nb = np / NB;
    const size_t c = np, mat = c * c;
    int rc;
    double* Lm = b->Lm + s0 * mat; double* Li = b->Li + s0 * mat;
    rescale_batched_kernel<<<dim3((np + 127) / 128, 1, Sa), 128, 0, st>>>(np, d, b->hyps + s0, b->Xraw, b->Xs + s0 * c * BO_MAX_DIM);
    BO_LAUNCH_CHECK(h);
    if ((rc = BO_DISPATCH_DP(b->dp, lml_launch_gram, h, b, s0, Sa, st))) return rc;
    BO_CUDA(h, cudaMemsetAsync(Li, 0, (size_t)Sa * mat * sizeof(double), st));
    for (int kb = 0; kb < nb; ++kb) {
        const size_t off = (size_t)kb * NB * ld + kb * NB;
        chol_panel_kernel<<<Sa * (nb - kb), 256, 0, st>>>(Lm + off, ld, Li + off, b->info + s0, kb * NB, mat, nb - kb, chol_panel_fused(), chol_panel_kpre(kb));
        BO_LAUNCH_CHECK(h);
        if (kb + 1 < nb && (rc = lml_gemm(h, b, kb, s0, Sa, st))) return rc;
    }
    leaf_inverse_kernel<<<Sa * nb, 256, 0, st>>>(Lm, ld, Li, mat, nb);
    BO_LAUNCH_CHECK(h);
    for (int li = nb - 1; li < (int)b->launches.size(); ++li)             // inverse levels, then K^-1 = L^-T L^-1
        if ((rc = lml_gemm(h, b, li, s0, Sa, st))) return rc;
    // alpha = L^-T (L^-1 r) + one refinement step against the factor (as in the single fit)
    const dim3 gv((np + 255) / 256, 1, Sa), gr(np / 8, 1, Sa), gt(np / 32, TRMVT_SPLITS, Sa);
    double* v1 = b->v1 + s0 * c; double* v2 = b->v2 + s0 * c; double* v3 = b->v3 + s0 * c; double* alpha = b->alpha + s0 * c;
    double* tpart = b->tpart + (size_t)s0 * TRMVT_SPLITS * c;
    resid_init_kernel<<<gv, 256, 0, st>>>(b->yv, n, np, mean, v1);
    trmv_lower_kernel<<<gr, 256, 0, st>>>(Li, ld, np, v1, v2, mat);
    trmv_lower_t_kernel<<<gt, 256, 0, st>>>(Li, ld, np, v2, tpart, mat);
    trmv_reduce_kernel<<<gv, 256, 0, st>>>(tpart, np, alpha, 0);
    trmv_lower_t_kernel<<<gt, 256, 0, st>>>(Lm, ld, np, alpha, tpart, mat);
    trmv_reduce_kernel<<<gv, 256, 0, st>>>(tpart, np, v2, 0);
    trmv_lower_kernel<<<gr, 256, 0, st>>>(Lm, ld, np, v2, v3, mat);
    sub_vec_kernel<<<gv, 256, 0, st>>>(v1, v3, np, v2);
    trmv_lower_kernel<<<gr, 256, 0, st>>>(Li, ld, np, v2, v3, mat);
    trmv_lower_t_kernel<<<gt, 256, 0, st>>>(Li, ld, np, v3, tpart, mat);
    trmv_reduce_kernel<<<gv, 256, 0, st>>>(tpart, np, alpha, 1);
    h->launches += 11;
    BO_CUDA(h, cudaGetLastError());
    return BO_DISPATCH_DP(b->dp, lml_launch_grad, h, b, s0, Sa, st);
}

int lml_impl(bo_handle* h, const double* X_dev, const double* y_dev, int n, int d, int kind, double mean,
             const double* theta_host, int R, double* lml_host, double* grad_host, int* status_host, cudaStream_t st) {
    if (!X_dev || !y_dev || !theta_host || !lml_host || !grad_host || !status_host || n < 1 || d < 1 || R < 1)
        return fail(h, BO_E_INVALID, "bo_lml_grad_batched: bad argument");
    if (d > BO_MAX_DIM) return fail(h, BO_E_CAPACITY, "bo_lml_grad_batched: d exceeds BO_MAX_DIM");
    if (kind != BO_KERNEL_MATERN52 && kind != BO_KERNEL_RBF && kind != BO_KERNEL_LINEAR_MATERN52)
        return fail(h, BO_E_INVALID, "bo_lml_grad_batched: unknown kernel kind");
    const int p = d + 2 + (kind == BO_KERNEL_LINEAR_MATERN52 ? 1 : 0);          // parameters per restart
    BO_CUDA(h, cudaSetDevice(h->device));
    const int np = round_up(n, PAD);
    // slots: as many restarts in lock step as ~6 GB of workspace allows (4 n^2 doubles per slot), at most 32
    const size_t per_slot = (size_t)np * np * 8 * 13 / 4;
    int S = (int)std::min<size_t>(32, std::max<size_t>(1, (6ull << 30) / per_slot));
    if (S > R) S = R;
    int rc = lml_prepare(h, n, d, S, st);
    if (rc) return rc;
    LmlBatch* b = static_cast<LmlBatch*>(h->lml_batch);
    S = b->S;                         // slots actually allocated (>= the request)
    Hyper zero{};   // inv_ls = 1, mean: staging of the unscaled inputs uses an identity scale
    for (int k = 0; k < BO_MAX_DIM; ++k) zero.inv_ls[k] = 1.0;
    stage_inputs_kernel<<<(np + 127) / 128, 128, 0, st>>>(X_dev, y_dev, n, d, np, zero, b->Xraw, b->Xs, b->yv);
    BO_LAUNCH_CHECK(h);

    for (int r0 = 0; r0 < R; r0 += S) {
        const int Sa = std::min(S, R - r0);
        for (int s = 0; s < Sa; ++s) {
            const double* th = theta_host + (size_t)(r0 + s) * p;
            Hyper& hy = b->hyps_host[s];
            hy.kind = kind; hy.d = d; hy.dp = b->dp; hy.mean = mean; hy.jitter = 0.0;
            hy.lin_v = kind == BO_KERNEL_LINEAR_MATERN52 ? exp(th[d + 2]) : 0.0;
            for (int k = 0; k < BO_MAX_DIM; ++k) {
                hy.inv_ls[k] = k < d ? exp(-th[k]) : 0.0;
                hy.lin_w[k] = k < d ? hy.lin_v * exp(2.0 * th[k]) : 0.0;
            }
            hy.outputscale = exp(th[d]); hy.noise = exp(th[d + 1]);
        }
        BO_CUDA(h, cudaMemcpyAsync(b->hyps, b->hyps_host, Sa * sizeof(Hyper), cudaMemcpyHostToDevice, st));
        BO_CUDA(h, cudaMemsetAsync(b->info, 0, Sa * sizeof(int), st));
        // slot groups on their own streams (fork after the parameter upload, join before the read-back)
        // measured (tools/lml_groups_probe.py): 4 groups win 8-10 % from 8 slots up, 2 groups 4 % at 4-7 slots of a large
        // problem; below that the split only shrinks the launches
        int G = Sa >= 8 ? 4 : (Sa >= 4 && np >= 2048) ? 2 : 1;
        { const char* ge = getenv("BO_B200_LML_GROUPS"); if (ge && atoi(ge) >= 1 && atoi(ge) <= LmlBatch::MAX_GROUPS && atoi(ge) <= Sa) G = atoi(ge); }
        BO_CUDA(h, cudaEventRecord(b->fork_ev, st));
        for (int g = 0, s0 = 0; g < G; ++g) {
            const int Sg = Sa / G + (g < Sa % G ? 1 : 0);
            BO_CUDA(h, cudaStreamWaitEvent(b->gstream[g], b->fork_ev, 0));
            if ((rc = lml_run_slots(h, b, s0, Sg, n, d, mean, b->gstream[g]))) return rc;
            BO_CUDA(h, cudaEventRecord(b->join_ev[g], b->gstream[g]));
            BO_CUDA(h, cudaStreamWaitEvent(st, b->join_ev[g], 0));
            s0 += Sg;
        }
        BO_CUDA(h, cudaMemcpyAsync(b->out_host, b->out, (size_t)Sa * (BO_MAX_DIM + 4) * sizeof(double), cudaMemcpyDeviceToHost, st));
        BO_CUDA(h, cudaMemcpyAsync(b->info_host, b->info, Sa * sizeof(int), cudaMemcpyDeviceToHost, st));
        BO_CUDA(h, cudaStreamSynchronize(st));
        for (int s = 0; s < Sa; ++s) {
            const int r = r0 + s;
            const int info = b->info_host[s];
            status_host[r] = info > n ? n : info;
            const double* o = b->out_host + (size_t)s * (BO_MAX_DIM + 4);
            if (info != 0) {
                lml_host[r] = -INFINITY;
                for (int k = 0; k < p; ++k) grad_host[(size_t)r * p + k] = 0.0;
            } else {
                lml_host[r] = o[0];
                for (int k = 0; k < p; ++k) grad_host[(size_t)r * p + k] = o[1 + k];
            }
        }
    }
    return 0;
}
