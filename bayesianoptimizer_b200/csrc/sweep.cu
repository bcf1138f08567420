// K4: fused acquisition sweep.
//
// One persistent CTA per SM.  Per block of SW_BN = 128 candidates a CTA
//   phase A  generates the candidates (in-kernel scrambled Sobol or an explicit pool), builds the
//            cross-covariance panel K(X, X*) in DMMA B-fragment order (its own L2/HBM-resident slot)
//            and the posterior mean k*^T alpha on the way;
//   phase B  contracts the packed lower-triangular L^-1 (A operand) with the panel on the FP64
//            tensor path (DMMA.8x8x4), both operands streamed by TMA bulk copies (UBLKCP) into a
//            3-stage shared-memory ring guarded by full/empty mbarriers (issued by one elected thread); each finished 128-row slab is squared and reduced per candidate in registers;
//   epilogue var = max(s2 - ||L^-1 k*||^2, min_var), EI / LogEI / UCB with erfc/erfcx in registers,
//            optional per-candidate outputs, CTA-local top-k by (value desc, index asc).
// A single-block merge kernel reduces the per-CTA lists.  Replaces the chunked pool scan + CPU topk
// of optimization/Bayesian7.py:664-682 and optimize_acqf's raw-sample scoring (Bayesian.py:105-112).
#include "common.cuh"
#include <cstdlib>

namespace bo {

constexpr long long IDX_EMPTY = 0x7fffffffffffffffLL;
constexpr int SW_MAX_SEG = 16;

struct SweepArgs {
    const double* Xs; const double* alpha; const double* Lp; double* panel;
    const double* cand; const bo_sobol* sobol;
    long long first_index, N, nblocks;
    int n, np, d;
    Hyper hyp;
    int acq; double best_f, sqrt_beta, min_var;
    int topk; double* part_val; long long* part_idx;
    double* mean_out; double* var_out; double* acq_out;
    int flags;      // bit1: skip the all-zero 8x8 blocks of diagonal tiles (default on)
    // small pools: every candidate block is split into G row segments (work items) so that all SMs are busy;
    // segment s covers row blocks [seg[s], seg[s+1]) of L^-1, per-row-block column sums meet in global memory and the
    // last-arriving CTA of a block runs the epilogue, adding them in ascending row-block order -- the same two
    // chains the unsplit path keeps in registers, so the result is bit-identical for every split
    int G; int seg[SW_MAX_SEG + 1];
    double* part_cs; double* part_mu; double* part_kss; int* counters;
    // SVGP predictive mode (SV kernels): second triangular factor J Ls^T J (packed like Lp), the panel that receives the
    // row-reversed interp term u = L^-1 k*, and the constant added to the prior variance (K_uu jitter + likelihood noise)
    const double* Lp2; double* panel2; double sv_add;
    // re-score pass of the sliced sweep's accuracy guard: candidate li of this launch is pool entry idx_map[li] (coordinates,
    // dense outputs and top-k index all go through the map); nullptr = the identity
    const long long* idx_map;
};
__device__ __forceinline__ long long pool_index(const SweepArgs& a, long long li) { return a.idx_map ? a.idx_map[li] : li; }

// ---- analytic acquisition (botorch.acquisition.analytic semantics) ------
__device__ __forceinline__ double log1mexp_d(double x) {
    return (x > -0.69314718055994530942) ? log(-expm1(x)) : log1p(-exp(x));
}

__device__ double acq_value(int kind, double mu, double var, double best_f, double sqrt_beta) {
    if (kind == BO_ACQ_VAR) return var;
    if (kind == BO_ACQ_MEAN) return mu;
    const double sigma = sqrt(var);
    if (kind == BO_ACQ_UCB) return fma(sqrt_beta, sigma, mu);
    const double u = (mu - best_f) / sigma;
    const double inv_sqrt2 = 0.70710678118654752440, inv_sqrt_2pi = 0.39894228040143267794;
    if (kind == BO_ACQ_EI) {
        const double phi = inv_sqrt_2pi * exp(-0.5 * u * u);
        const double Phi = 0.5 * erfc(-u * inv_sqrt2);
        return sigma * fma(u, Phi, phi);
    }
    // LogEI
    double lh;
    if (u > -1.0) {
        const double phi = inv_sqrt_2pi * exp(-0.5 * u * u);
        const double Phi = 0.5 * erfc(-u * inv_sqrt2);
        lh = log(fma(u, Phi, phi));
    } else {
        const double w = log(erfcx(-u * inv_sqrt2) * fabs(u)) + 0.22579135264472743236;   // + log(pi/2)/2
        lh = -0.5 * u * u - 0.91893853320467274178 + log1mexp_d(w);                       // - log(2 pi)/2
    }
    return log(sigma) + lh;
}

// candidate coordinates (unscaled) of global index gi; point 0 reproduces torch's float32 first point
template <int DP>
__device__ __forceinline__ void sobol_point(const uint32_t* dirs /*[DP][30]*/, const uint32_t* shift, int d,
                                            long long gi, double* x) {
    const unsigned long long gray = (unsigned long long)gi ^ ((unsigned long long)gi >> 1);
    uint32_t acc[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) acc[k] = shift[k];
    for (int b = 0; b < BO_SOBOL_BITS; ++b) {
        if ((gray >> b) & 1ULL) {
#pragma unroll
            for (int k = 0; k < DP; ++k) acc[k] ^= dirs[k * BO_SOBOL_BITS + b];
        }
    }
#pragma unroll
    for (int k = 0; k < DP; ++k) {
        double v = (gi == 0) ? (double)(float)acc[k] : (double)acc[k];
        x[k] = (k < d) ? v * 9.31322574615478515625e-10 : 0.0;   // 2^-30
    }
}

// ---- shared-memory carve-up ------------------------------------------------------------------
struct SweepSmem {
    static constexpr int STAGE_BYTES = 2 * SW_TILE * 8;                 // A tile + B tile
    static constexpr int OFF_BAR   = SW_STAGES * STAGE_BYTES;           // full[S], empty[S]
    static constexpr int OFF_COL   = OFF_BAR + 64;                      // colsum[2][SW_BN]
    static constexpr int OFF_MU    = OFF_COL + 2 * SW_BN * 8;           // mu[SW_BN]
    static constexpr int OFF_KSS   = OFF_MU + SW_BN * 8;                // kss[SW_BN]: prior variance k(x*,x*) (linear + Matern kind)
    static constexpr int OFF_TKV   = OFF_KSS + SW_BN * 8;               // tk_val[64]
    static constexpr int OFF_TKI   = OFF_TKV + BO_MAX_TOPK * 8;         // tk_idx[64]
    static constexpr int OFF_ACQ   = OFF_TKI + BO_MAX_TOPK * 8;         // acq[SW_BN]
    static constexpr int OFF_CMASK = OFF_ACQ + SW_BN * 8;               // cmask[4], is_last
    static constexpr int OFF_SOB   = OFF_CMASK + 32;                    // dirs[16][30] + shift[16]
    static constexpr int BYTES     = OFF_SOB + (BO_MAX_DIM * BO_SOBOL_BITS + BO_MAX_DIM) * 4;
};

// SV = true: whitened-SVGP predictive variance k** + jitter - ||u||^2 + ||Ls^T u||^2 (+ noise), u = L^-1 k*
// (gpytorch VariationalStrategy as driven by optimization/Bayesian7.py:664-671): phase B runs twice over the same ring --
// pass 0 contracts L^-1 with the K panel and parks u, ROW-REVERSED, in a second panel (so that the upper-triangular
// Ls^T becomes the lower-triangular J Ls^T J the tile loop already knows how to stream); pass 1 contracts that with u.
template <int DP, int KIND, bool SV>
__global__ void __launch_bounds__(SW_THREADS, 1) sweep_kernel(const SweepArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full  = reinterpret_cast<uint64_t*>(smem + SweepSmem::OFF_BAR);
    uint64_t* empty = full + SW_STAGES;
    double* colsum  = reinterpret_cast<double*>(smem + SweepSmem::OFF_COL);
    double* mu_s    = reinterpret_cast<double*>(smem + SweepSmem::OFF_MU);
    double* kss_s   = reinterpret_cast<double*>(smem + SweepSmem::OFF_KSS);
    double* tkv     = reinterpret_cast<double*>(smem + SweepSmem::OFF_TKV);
    long long* tki  = reinterpret_cast<long long*>(smem + SweepSmem::OFF_TKI);
    double* acq_s   = reinterpret_cast<double*>(smem + SweepSmem::OFF_ACQ);
    unsigned* cmask = reinterpret_cast<unsigned*>(smem + SweepSmem::OFF_CMASK);
    int* is_last    = reinterpret_cast<int*>(smem + SweepSmem::OFF_CMASK + 16);
    uint32_t* dirs  = reinterpret_cast<uint32_t*>(smem + SweepSmem::OFF_SOB);
    uint32_t* shift = dirs + BO_MAX_DIM * BO_SOBOL_BITS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int nbm = a.np / SW_BM;
    constexpr int KCH = SW_BM / SW_BK;
    double* panel = a.panel + (size_t)blockIdx.x * (a.np / SW_BK) * SW_TILE;
    double* panel2 = SV ? a.panel2 + (size_t)blockIdx.x * (a.np / SW_BK) * SW_TILE : nullptr;

    if (tid == 0) {
        for (int s = 0; s < SW_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SW_CONSUMER_WARPS); }
        fence_mbar_init();
    }
    if (tid < BO_MAX_TOPK) { tkv[tid] = -INFINITY; tki[tid] = IDX_EMPTY; }
    if (a.sobol) {
        for (int e = tid; e < DP * BO_SOBOL_BITS; e += SW_THREADS)
            dirs[e] = a.sobol->direction[e / BO_SOBOL_BITS][e % BO_SOBOL_BITS];
        if (tid < DP) shift[tid] = a.sobol->shift[tid];
    }
    __syncthreads();

    int stage = 0; uint32_t phase = 0;      // consumer ring position
    int pstage = 0; uint32_t pphase = 0;    // producer ring position (thread 0 issues the TMA bulk copies)

    for (long long w = blockIdx.x; w < a.nblocks * a.G; w += gridDim.x) {
        const long long blk = w / a.G;
        const int sg = (int)(w % a.G);
        const int ib0 = a.seg[sg], ib1 = a.seg[sg + 1];      // row blocks of L^-1 handled by this work item
        // ================= phase A: candidates, K(X, X*) panel, posterior mean =================
        {
            double xc[2][DP];
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                long long li = blk * SW_BN + (warp + 8 * gi) * 8 + g;
                if (li >= a.N) li = a.N - 1;
                li = pool_index(a, li);
                if (a.cand) {
#pragma unroll
                    for (int k = 0; k < DP; ++k) xc[gi][k] = (k < a.d) ? a.cand[(size_t)li * a.d + k] : 0.0;
                } else {
                    sobol_point<DP>(dirs, shift, a.d, a.first_index + li, xc[gi]);
                }
#pragma unroll
                for (int k = 0; k < DP; ++k) xc[gi][k] = __dmul_rn(xc[gi][k], a.hyp.inv_ls[k]);     // (never contracted into the differences below: every code instance must round alike)
            }
            // linear + Matern kind: weighted copies for the inner-product term and the candidates' prior variance
            double xw[2][KIND == BO_KERNEL_LINEAR_MATERN52 ? DP : 1];
            if (KIND == BO_KERNEL_LINEAR_MATERN52) {
#pragma unroll
                for (int gi = 0; gi < 2; ++gi) {
                    double nn = 0.0;
#pragma unroll
                    for (int k = 0; k < DP; ++k) { xw[gi][k] = a.hyp.lin_w[k] * xc[gi][k]; nn = fma(xw[gi][k], xc[gi][k], nn); }
                    if (q == 0) kss_s[(warp + 8 * gi) * 8 + g] = a.hyp.outputscale * (nn + 1.0);
                }
            }
            double mu0 = 0.0, mu1 = 0.0;
            const int nj8 = ib1 * (SW_BM / 8);               // panel rows this segment contracts over
            // The scaled observations X~ and alpha are staged through the (idle) stage buffers in double-buffered
            // chunks with cp.async, so the panel build reads them from shared memory instead of stalling on L2.
            // PA_SL 8-row slices per iteration: 4 * PA_SL independent kernel evaluations per lane in flight.
            constexpr int PA_SL = (KIND == BO_KERNEL_LINEAR_MATERN52 && DP >= 12) ? 2 : 4, PA_R = 2 * PA_SL;   // (xw doubles the candidate registers)
            constexpr int XCH = (DP <= 8) ? 1024 : 512;          // rows per chunk
            constexpr int XP = DP + 2;                           // padded row pitch (doubles): conflict-free LDS.128 over 4 rows
            constexpr int XBUF = XCH * XP + XCH;                 // doubles per buffer: rows + alpha
            static_assert(2 * XBUF * 8 <= SW_STAGES * SweepSmem::STAGE_BYTES, "X~ staging must fit into the stage buffers");
            double* xstage = reinterpret_cast<double*>(smem);
            const int nrows = nj8 * 8;
            const int nchunks = (nrows + XCH - 1) / XCH;
            auto load_chunk = [&](int c) {
                double* xb = xstage + (c & 1) * XBUF;
                double* ab = xb + XCH * XP;
                const int r0 = c * XCH, rows = min(XCH, nrows - r0);
                for (int e = tid; e < rows * (DP / 2); e += SW_THREADS) {
                    const int r = e / (DP / 2), k = e % (DP / 2);
                    cp_async16(xb + r * XP + 2 * k, a.Xs + (size_t)(r0 + r) * BO_MAX_DIM + 2 * k);
                }
                for (int e = tid; e < rows / 2; e += SW_THREADS) cp_async16(ab + 2 * e, a.alpha + r0 + 2 * e);
                cp_async_commit();
            };
            load_chunk(0);
            for (int c = 0; c < nchunks; ++c) {
                if (c + 1 < nchunks) { load_chunk(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
                __syncthreads();
                const double* xb = xstage + (c & 1) * XBUF;
                const double* ab = xb + XCH * XP;
                const int j8_end = min(nj8, (c + 1) * (XCH / 8));
                for (int j8 = c * (XCH / 8); j8 < j8_end; j8 += PA_SL) {
                    double x[PA_R][DP];                               // rows j8*8 + q + 4r
                    double al[PA_R];
#pragma unroll
                    for (int r = 0; r < PA_R; ++r) {
                        const int jr = j8 * 8 + q + 4 * r - c * XCH;
                        const double2* row = reinterpret_cast<const double2*>(xb + jr * XP);
#pragma unroll
                        for (int k = 0; k < DP / 2; ++k) {
                            const double2 t = row[k];
                            x[r][2 * k] = t.x; x[r][2 * k + 1] = t.y;
                        }
                        al[r] = ab[jr];
                    }
                    double kv[2][PA_R];
#pragma unroll
                    for (int gi = 0; gi < 2; ++gi)
#pragma unroll
                        for (int r = 0; r < PA_R; ++r) {
                            double sq = 0.0, lin = 0.0;
#pragma unroll
                            for (int k = 0; k < DP; ++k) {
                                const double df = __dsub_rn(xc[gi][k], x[r][k]);
                                sq = fma(df, df, sq);
                                if (KIND == BO_KERNEL_LINEAR_MATERN52) lin = fma(xw[gi][k], x[r][k], lin);
                            }
                            const double v = kernel_pair_t<KIND>(sq, lin, a.hyp.outputscale);
                            kv[gi][r] = (j8 * 8 + q + 4 * r < a.n) ? v : 0.0;
                        }
#pragma unroll
                    for (int r = 0; r < PA_R; ++r) { mu0 = fma(kv[0][r], al[r], mu0); mu1 = fma(kv[1][r], al[r], mu1); }
#pragma unroll
                    for (int gi = 0; gi < 2; ++gi)
#pragma unroll
                        for (int hh = 0; hh < PA_SL; ++hh) {
                            const int jj = j8 + hh;
                            double* dst = panel + (size_t)(jj / (SW_BK / 8)) * SW_TILE + (((warp + 8 * gi) * (SW_BK / 8) + (jj % (SW_BK / 8))) * 64 + lane * 2);
                            *reinterpret_cast<double2*>(dst) = make_double2(kv[gi][2 * hh], kv[gi][2 * hh + 1]);
                        }
                }
                __syncthreads();          // every warp is done with this buffer before chunk c + 2 overwrites it
            }
            mu0 += __shfl_xor_sync(0xffffffffu, mu0, 1); mu0 += __shfl_xor_sync(0xffffffffu, mu0, 2);
            mu1 += __shfl_xor_sync(0xffffffffu, mu1, 1); mu1 += __shfl_xor_sync(0xffffffffu, mu1, 2);
            if (q == 0) { mu_s[warp * 8 + g] = mu0; mu_s[(warp + 8) * 8 + g] = mu1; }
            __threadfence();
            fence_proxy_async();      // generic-proxy panel writes -> visible to the async-proxy (TMA) reads
        }
        __syncthreads();

        // ================= phase B: ||L^-1 k*||^2 on the DMMA path ============================
        {
            const int wm = warp >> 2, wn = warp & 3;
            double colsq[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) colsq[e] = 0.0;
          for (int pass = 0; pass < (SV ? 2 : 1); ++pass) {
            const double* Asrc = (SV && pass) ? a.Lp2 : a.Lp;
            const double* Bsrc = (SV && pass) ? panel2 : panel;
            if (SV && pass) {             // u is complete: generic-proxy panel2 writes -> visible to the TMA reads
                __threadfence();
                fence_proxy_async();
                __syncthreads();
            }
            const long long T = ((long long)ib1 * (ib1 + 1) / 2 - (long long)ib0 * (ib0 + 1) / 2) * KCH;   // pipeline stages of this item
            int pib = ib0, pkc = 0;                                          // producer position (thread 0)
            long long issued = 0;
            auto issue = [&]() {
                mbar_wait(&empty[pstage], pphase ^ 1);
                unsigned char* sb = smem + pstage * SweepSmem::STAGE_BYTES;
                const double* At = Asrc + ((size_t)pib * (pib + 1) / 2 * KCH + pkc) * SW_TILE;
                mbar_expect_tx(&full[pstage], SweepSmem::STAGE_BYTES);
                bulk_g2s(sb, At, SW_TILE * 8, &full[pstage]);
                bulk_g2s(sb + SW_TILE * 8, Bsrc + (size_t)pkc * SW_TILE, SW_TILE * 8, &full[pstage]);
                if (++pstage == SW_STAGES) { pstage = 0; pphase ^= 1; }
                if (++pkc == (pib + 1) * KCH) { pkc = 0; ++pib; }
                ++issued;
            };
            if (tid == 0)
                while (issued < T && issued < SW_STAGES - 1) issue();

            bool ready = false;              // full[stage] already observed complete by the early probe
            for (int ib = ib0; ib < ib1; ++ib) {
                double acc[8][4][2];
#pragma unroll
                for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                const int nkc = (ib + 1) * KCH;
                for (int kc = 0; kc < nkc; ++kc) {
                    // refill the slot released one iteration ago (prefetch distance SW_STAGES - 1)
                    if (tid == 0 && issued < T) issue();
                    __syncwarp();
                    if (!ready) mbar_wait(&full[stage], phase);       // usually already probed half a stage ago
                    const int nstage = (stage + 1 == SW_STAGES) ? 0 : stage + 1;
                    const uint32_t nphase = (nstage == 0) ? phase ^ 1 : phase;
                    const double* As = reinterpret_cast<const double*>(smem + stage * SweepSmem::STAGE_BYTES);
                    const double* Bs = As + SW_TILE;
                    const int kdiag = kc - ib * KCH;                      // >= 0 inside the diagonal tile
                    if (kdiag < 0 || !(a.flags & 2)) {
#pragma unroll
                        for (int k8 = 0; k8 < SW_BK / 8; ++k8) {
                            double2 b[4];
#pragma unroll
                            for (int ni = 0; ni < 4; ++ni)
                                b[ni] = *reinterpret_cast<const double2*>(Bs + (((wn * 4 + ni) * (SW_BK / 8) + k8) * 64 + lane * 2));
                            // probe the next stage's barrier in the middle of this one: the ~90-cycle try_wait hides
                            // behind the DMMA stream instead of idling the pipe at every stage boundary
                            if (k8 == (SW_BK / 8) / 2) ready = mbar_try_wait(&full[nstage], nphase);
#pragma unroll
                            for (int mi = 0; mi < 8; ++mi) {
                                const double2 av = *reinterpret_cast<const double2*>(As + (((wm * 8 + mi) * (SW_BK / 8) + k8) * 64 + lane * 2));
#pragma unroll
                                for (int ni = 0; ni < 4; ++ni) {
                                    dmma884(acc[mi][ni][0], acc[mi][ni][1], av.x, b[ni].x);
                                    dmma884(acc[mi][ni][0], acc[mi][ni][1], av.y, b[ni].y);
                                }
                            }
                        }
                    } else {
                        ready = false;
                        // diagonal tile: 8x8 blocks strictly above the diagonal hold zeros -> skip their DMMAs
#pragma unroll
                        for (int k8 = 0; k8 < SW_BK / 8; ++k8) {
                            const int mi_min = kdiag * (SW_BK / 8) + k8 - wm * 8;     // first row block with data
                            if (mi_min < 8) {
                                double2 b[4];
#pragma unroll
                                for (int ni = 0; ni < 4; ++ni)
                                    b[ni] = *reinterpret_cast<const double2*>(Bs + (((wn * 4 + ni) * (SW_BK / 8) + k8) * 64 + lane * 2));
#pragma unroll
                                for (int mi = 0; mi < 8; ++mi) {
                                    if (mi >= mi_min) {
                                        const double2 av = *reinterpret_cast<const double2*>(As + (((wm * 8 + mi) * (SW_BK / 8) + k8) * 64 + lane * 2));
#pragma unroll
                                        for (int ni = 0; ni < 4; ++ni) {
                                            dmma884(acc[mi][ni][0], acc[mi][ni][1], av.x, b[ni].x);
                                            dmma884(acc[mi][ni][0], acc[mi][ni][1], av.y, b[ni].y);
                                        }
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);
                    if (++stage == SW_STAGES) { stage = 0; phase ^= 1; }
                }
                if (SV && pass == 0) {
                    // park u = (L^-1 k*) rows of this block in panel2, row-reversed (i -> np-1-i), in B-fragment order
                    double* t2 = panel2 + ((size_t)(nbm - 1 - ib) * KCH + (1 - wm) * 2) * SW_TILE;
#pragma unroll
                    for (int mi = 0; mi < 8; ++mi) {
                        double* tm = t2 + ((7 - mi) >> 2) * SW_TILE + ((7 - mi) & 3) * 64 + (3 - (g & 3)) * 2 + (1 - (g >> 2));
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                            for (int e2 = 0; e2 < 2; ++e2)
                                tm[((wn * 4 + ni) * (SW_BK / 8)) * 64 + (2 * q + e2) * 8] = acc[mi][ni][e2];
                    }
                }
                // canonical reduction (independent of the segment split): per row block, square-sum this thread's 8
                // row groups, butterfly over the 8 row lanes, then add the row block's total to the running sum in
                // ascending row-block order -- or park it in global memory for the finaliser to add in that order
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    double t = 0.0;
#pragma unroll
                    for (int mi = 0; mi < 8; ++mi) t = fma(acc[mi][e >> 1][e & 1], acc[mi][e >> 1][e & 1], t);
                    t += __shfl_xor_sync(0xffffffffu, t, 4);
                    t += __shfl_xor_sync(0xffffffffu, t, 8);
                    t += __shfl_xor_sync(0xffffffffu, t, 16);
                    if (a.G > 1) {
                        if (g == 0)
                            a.part_cs[(((size_t)blk * nbm + ib) * 2 + wm) * SW_BN + wn * 32 + (e >> 1) * 8 + 2 * q + (e & 1)] = t;
                    } else {
                        colsq[e] += (SV && pass) ? -t : t;      // SV: ss = ||u||^2 - ||Ls^T u||^2
                    }
                }
            }
          }   // pass
            if (g == 0) {
#pragma unroll
                for (int e = 0; e < 8; ++e) colsum[wm * SW_BN + wn * 32 + (e >> 1) * 8 + 2 * q + (e & 1)] = colsq[e];
            }
        }
        __syncthreads();

        // ================= split blocks: publish partials, last arriver finalises =================
        if (a.G > 1) {
            if (tid < SW_BN && ib1 == nbm) {                 // the last segment saw every row
                a.part_mu[(size_t)blk * SW_BN + tid] = mu_s[tid];
                if (KIND == BO_KERNEL_LINEAR_MATERN52) a.part_kss[(size_t)blk * SW_BN + tid] = kss_s[tid];
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) *is_last = (atomicAdd(&a.counters[blk], 1) == a.G - 1) ? 1 : 0;
            __syncthreads();
            if (!*is_last) continue;                         // uniform: every thread reads the same flag
            __threadfence();
        }
        // ================= epilogue: variance, acquisition, CTA-local top-k =====================
        if (tid < SW_BN) {
            const long long li = blk * SW_BN + tid;
            double ss, mu_c, prior = a.hyp.outputscale;
            if (KIND == BO_KERNEL_LINEAR_MATERN52)
                prior = (a.G > 1) ? __ldcg(a.part_kss + (size_t)blk * SW_BN + tid) : kss_s[tid];
            if (a.G > 1) {
                double r0 = 0.0, r1 = 0.0;                   // the same two ascending chains the unsplit path keeps in registers
                for (int ib = 0; ib < nbm; ++ib) {
                    r0 += __ldcg(a.part_cs + (((size_t)blk * nbm + ib) * 2 + 0) * SW_BN + tid);
                    r1 += __ldcg(a.part_cs + (((size_t)blk * nbm + ib) * 2 + 1) * SW_BN + tid);
                }
                ss = r0 + r1;
                mu_c = __ldcg(a.part_mu + (size_t)blk * SW_BN + tid);
            } else {
                ss = colsum[tid] + colsum[SW_BN + tid];
                mu_c = mu_s[tid];
            }
            const double var = fmax((SV ? prior + a.sv_add : prior) - ss, a.min_var);
            const double mean = a.hyp.mean + mu_c;
            double v = acq_value(a.acq, mean, var, a.best_f, a.sqrt_beta);
            const long long pi = li < a.N ? pool_index(a, li) : li;
            if (li < a.N) {
                if (a.mean_out) a.mean_out[pi] = mean;
                if (a.var_out) a.var_out[pi] = var;
                if (a.acq_out) a.acq_out[pi] = v;
            }
            if (!(v == v)) v = -INFINITY;                    // NaN ranks last
            acq_s[tid] = v;
            // only candidates that beat the current k-th entry can enter the list (the list only improves)
            bool beats = false;
            if (a.topk > 0 && li < a.N) beats = tk_better(v, a.first_index + pi, tkv[a.topk - 1], tki[a.topk - 1]);
            const unsigned m = __ballot_sync(0xffffffffu, beats);
            if (lane == 0) cmask[warp] = m;
        }
        __syncthreads();
        if (tid == 0 && a.topk > 0) {
            const int K = a.topk;
            for (int w = 0; w < SW_BN / 32; ++w) {
                unsigned m = cmask[w];
                while (m) {
                    const int c = w * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    const double v = acq_s[c];
                    const long long gi = a.first_index + pool_index(a, blk * SW_BN + c);
                    if (!tk_better(v, gi, tkv[K - 1], tki[K - 1])) continue;
                    int p = K - 1;
                    while (p > 0 && tk_better(v, gi, tkv[p - 1], tki[p - 1])) { tkv[p] = tkv[p - 1]; tki[p] = tki[p - 1]; --p; }
                    tkv[p] = v; tki[p] = gi;
                }
            }
        }
        // no barrier here: the next block's phase A touches none of tkv/tki/acq_s/cmask, and two barriers
        // separate this insertion from the next epilogue
    }
    __syncthreads();
    if (tid < BO_MAX_TOPK && a.part_val) {
        a.part_val[(size_t)blockIdx.x * BO_MAX_TOPK + tid] = tkv[tid];
        a.part_idx[(size_t)blockIdx.x * BO_MAX_TOPK + tid] = tki[tid];
    }
}

__global__ void topk_merge_kernel(double* __restrict__ pv, long long* __restrict__ pi, int total, int topk, double* __restrict__ vals,
                                  long long* __restrict__ idx);

// ---- merge of the per-CTA top-k lists (k rounds of a block-wide arg-best) ------------------
__global__ void __launch_bounds__(1024) topk_merge_kernel(double* __restrict__ pv, long long* __restrict__ pi,
                                                          int total, int topk, double* __restrict__ vals,
                                                          long long* __restrict__ idx) {
    __shared__ double sv[32];
    __shared__ long long si[32];
    __shared__ int sp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int r = 0; r < topk; ++r) {
        double bv = -INFINITY; long long bi = IDX_EMPTY; int bp = -1;
        for (int e = tid; e < total; e += 1024) {
            const long long ii = pi[e];
            if (ii == IDX_EMPTY) continue;
            const double v = pv[e];
            if (bp < 0 || tk_better(v, ii, bv, bi)) { bv = v; bi = ii; bp = e; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            int op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (op >= 0 && (bp < 0 || tk_better(ov, oi, bv, bi))) { bv = ov; bi = oi; bp = op; }
        }
        if (lane == 0) { sv[warp] = bv; si[warp] = bi; sp[warp] = bp; }
        __syncthreads();
        if (warp == 0) {
            bv = sv[lane]; bi = si[lane]; bp = sp[lane];
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
                int op = __shfl_xor_sync(0xffffffffu, bp, o);
                if (op >= 0 && (bp < 0 || tk_better(ov, oi, bv, bi))) { bv = ov; bi = oi; bp = op; }
            }
            if (lane == 0) {
                vals[r] = (bp >= 0) ? bv : -INFINITY;
                idx[r] = (bp >= 0) ? bi : -1;
                if (bp >= 0) pi[bp] = IDX_EMPTY;
            }
        }
        __syncthreads();
    }
}

// ---- independent slow path (triage / tests): plain loads, row-major L^-1, no TMA, no DMMA ----
template <int DP>
__global__ void __launch_bounds__(256) sweep_reference_kernel(const SweepArgs a, const double* __restrict__ Li, int ld) {
    extern __shared__ double ks[];           // k*[np]
    __shared__ double red[8];
    __shared__ double xs_c[DP];
    const long long li = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        double x[DP];
        if (a.cand) { for (int k = 0; k < DP; ++k) x[k] = (k < a.d) ? a.cand[(size_t)li * a.d + k] : 0.0; }
        else {
            const uint32_t* dirs = &a.sobol->direction[0][0];      // straight from global memory
            const unsigned long long gi = (unsigned long long)(a.first_index + li);
            const unsigned long long gray = gi ^ (gi >> 1);
            for (int k = 0; k < DP; ++k) {
                uint32_t acc = a.sobol->shift[k];
                for (int b = 0; b < BO_SOBOL_BITS; ++b) if ((gray >> b) & 1ULL) acc ^= dirs[k * BO_SOBOL_BITS + b];
                double v = (gi == 0) ? (double)(float)acc : (double)acc;
                x[k] = (k < a.d) ? v * 9.31322574615478515625e-10 : 0.0;
            }
        }
        for (int k = 0; k < DP; ++k) xs_c[k] = x[k] * a.hyp.inv_ls[k];
    }
    __syncthreads();
    double mu = 0.0;
    for (int j = tid; j < a.np; j += 256) {
        double v = 0.0;
        if (j < a.n) {
            double sq = 0.0, lin = 0.0;
            for (int k = 0; k < DP; ++k) {
                const double xj = a.Xs[(size_t)j * BO_MAX_DIM + k];
                const double df = xs_c[k] - xj;
                sq = fma(df, df, sq);
                lin = fma(a.hyp.lin_w[k] * xs_c[k], xj, lin);
            }
            v = kernel_value(a.hyp.kind, sq, a.hyp.outputscale);
            if (a.hyp.kind == BO_KERNEL_LINEAR_MATERN52) v = fma(a.hyp.outputscale, lin, v);
            mu = fma(v, a.alpha[j], mu);
        }
        ks[j] = v;
    }
    __syncthreads();
    double ss = 0.0;
    for (int i = warp; i < a.n; i += 8) {
        const double* row = Li + (size_t)i * ld;
        double s = 0.0;
        for (int j = lane; j <= i; j += 32) s = fma(row[j], ks[j], s);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        ss = fma(s, s, ss);        // identical in every lane
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mu += __shfl_xor_sync(0xffffffffu, mu, o);
    __syncthreads();
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    double sst = 0.0;
    for (int w = 0; w < 8; ++w) sst += red[w];
    __syncthreads();
    if (lane == 0) red[warp] = mu;
    __syncthreads();
    if (tid == 0) {
        double m = 0.0;
        for (int w = 0; w < 8; ++w) m += red[w];
        double prior = a.hyp.outputscale;
        if (a.hyp.kind == BO_KERNEL_LINEAR_MATERN52) {
            double nn = 0.0;
            for (int k = 0; k < DP; ++k) nn = fma(a.hyp.lin_w[k] * xs_c[k], xs_c[k], nn);
            prior = a.hyp.outputscale * (nn + 1.0);
        }
        const double var = fmax(prior - sst, a.min_var);
        const double mean = a.hyp.mean + m;
        a.mean_out[li] = mean;
        a.var_out[li] = var;
        a.acq_out[li] = acq_value(a.acq, mean, var, a.best_f, a.sqrt_beta);
    }
}

// scan of a dense score array into per-block top-k lists (reference path only)
__global__ void __launch_bounds__(256) score_topk_kernel(const double* __restrict__ acq, long long N, long long first_index,
                                                         int topk, double* __restrict__ pv, long long* __restrict__ pi) {
    __shared__ double tkv[BO_MAX_TOPK];
    __shared__ long long tki[BO_MAX_TOPK];
    if (threadIdx.x < BO_MAX_TOPK) { tkv[threadIdx.x] = -INFINITY; tki[threadIdx.x] = IDX_EMPTY; }
    __syncthreads();
    if (threadIdx.x == 0 && topk > 0) {
        for (long long li = blockIdx.x; li < N; li += gridDim.x) {
            double v = acq[li]; if (!(v == v)) v = -INFINITY;
            const long long gi = first_index + li;
            if (!tk_better(v, gi, tkv[topk - 1], tki[topk - 1])) continue;
            int p = topk - 1;
            while (p > 0 && tk_better(v, gi, tkv[p - 1], tki[p - 1])) { tkv[p] = tkv[p - 1]; tki[p] = tki[p - 1]; --p; }
            tkv[p] = v; tki[p] = gi;
        }
    }
    __syncthreads();
    if (threadIdx.x < BO_MAX_TOPK) {
        pv[(size_t)blockIdx.x * BO_MAX_TOPK + threadIdx.x] = tkv[threadIdx.x];
        pi[(size_t)blockIdx.x * BO_MAX_TOPK + threadIdx.x] = tki[threadIdx.x];
    }
}

template <int DP>
__global__ void sobol_points_kernel(const bo_sobol* __restrict__ sob, const long long* __restrict__ idx, long long N, int d,
                                    double* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double x[DP];
    sobol_point<DP>(&sob->direction[0][0], sob->shift, d, idx[i], x);
    for (int k = 0; k < d; ++k) out[i * d + k] = x[k];
}

// ---- host side ----------------------------------------------------------------------------------
// Choose the row-segment count G for a pool of `nblocks` candidate blocks: with fewer blocks than a few waves
// of SMs, whole-block work items leave SMs idle (10^4 candidates = 79 blocks on 148 SMs), so blocks are split into
// G stage-balanced row segments.  Cost model: waves(nblocks * G) * (1/G + panel build) + a small per-segment penalty.
static int choose_segments(int sm, long long nblocks, int nbm, int* seg) {
    int best = 1; double best_cost = 1e300;
    const int gmax = nbm < SW_MAX_SEG ? nbm : SW_MAX_SEG;
    if (nblocks >= 4LL * sm) { seg[0] = 0; seg[1] = nbm; return 1; }
    for (int G = 1; G <= gmax; ++G) {
        const double waves = (double)((nblocks * G + sm - 1) / sm);
        const double cost = waves * (1.0 / G + 0.04) + 0.002 * G;   // the last segment rebuilds the whole panel (~4% of a block)
        if (cost < best_cost - 1e-12) { best_cost = cost; best = G; }
    }
    // stage-balanced boundaries: stages up to row block b  ~  b (b + 1) / 2
    const double total = 0.5 * nbm * (nbm + 1.0);
    seg[0] = 0;
    for (int s2 = 1; s2 < best; ++s2) {
        const double target = total * s2 / best;
        int b = (int)floor((-1.0 + sqrt(1.0 + 8.0 * target)) / 2.0 + 0.5);
        if (b <= seg[s2 - 1]) b = seg[s2 - 1] + 1;
        if (b > nbm - (best - s2)) b = nbm - (best - s2);
        seg[s2] = b;
    }
    seg[best] = nbm;
    return best;
}

static int ensure_split_ws(bo_handle* h, long long nblocks, int G, int nbm) {
    (void)G;
    const size_t need = ((size_t)nblocks * nbm * 2 * SW_BN + 2 * (size_t)nblocks * SW_BN) * sizeof(double) + (size_t)nblocks * sizeof(int) + 256;
    if (need > h->split_bytes) {
        if (h->split_ws) cudaFree(h->split_ws);
        h->split_ws = nullptr; h->split_bytes = 0;
        BO_CUDA(h, cudaMalloc(&h->split_ws, need));
        h->split_bytes = need;
    }
    return 0;
}

static int ensure_sweep_ws(bo_handle* h, int grid, bool fp64_panel = true, int lists = 0) {
    // the sliced sweep keeps its own int8 panels: it only needs the per-CTA top-k lists from here (`lists` of them when
    // more than `grid`: its re-score pass appends the FP64 kernel's lists behind its own)
    const size_t need = fp64_panel ? (size_t)grid * (h->np / SW_BK) * SW_TILE * sizeof(double) : 0;
    if (lists > grid) grid = lists;
    if (need > h->panel_bytes) {
        if (h->panel) cudaFree(h->panel);
        h->panel = nullptr; h->panel_bytes = 0;
        BO_CUDA(h, cudaMalloc(&h->panel, need));
        h->panel_bytes = need;
    }
    if (h->svgp && need > h->panel2_bytes) {
        if (h->panel2) cudaFree(h->panel2);
        h->panel2 = nullptr; h->panel2_bytes = 0;
        BO_CUDA(h, cudaMalloc(&h->panel2, need));
        h->panel2_bytes = need;
    }
    if (grid > h->part_grid) {
        if (h->part_val) cudaFree(h->part_val);
        if (h->part_idx) cudaFree(h->part_idx);
        h->part_val = nullptr; h->part_idx = nullptr; h->part_grid = 0;
        BO_CUDA(h, cudaMalloc(&h->part_val, (size_t)grid * BO_MAX_TOPK * sizeof(double)));
        BO_CUDA(h, cudaMalloc(&h->part_idx, (size_t)grid * BO_MAX_TOPK * sizeof(int64_t)));
        h->part_grid = grid;
    }
    return 0;
}

template <int DP, int KIND, bool SV>
static int launch_sweep_k(bo_handle* h, const SweepArgs& a, int grid, cudaStream_t st) {
    // the opt-in is per device context: set it on every launch (a process may hold handles on several GPUs)
    BO_CUDA(h, cudaFuncSetAttribute(sweep_kernel<DP, KIND, SV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SweepSmem::BYTES));
    sweep_kernel<DP, KIND, SV><<<grid, SW_THREADS, SweepSmem::BYTES, st>>>(a);
    BO_LAUNCH_CHECK(h);
    return 0;
}
template <int DP>
static int launch_sweep(bo_handle* h, const SweepArgs& a, int grid, cudaStream_t st) {
    const bool sv = a.Lp2 != nullptr;
    switch (a.hyp.kind) {
        case BO_KERNEL_MATERN52:
            return sv ? launch_sweep_k<DP, BO_KERNEL_MATERN52, true>(h, a, grid, st) : launch_sweep_k<DP, BO_KERNEL_MATERN52, false>(h, a, grid, st);
        case BO_KERNEL_LINEAR_MATERN52:
            return sv ? launch_sweep_k<DP, BO_KERNEL_LINEAR_MATERN52, true>(h, a, grid, st)
                      : launch_sweep_k<DP, BO_KERNEL_LINEAR_MATERN52, false>(h, a, grid, st);
        default:
            return sv ? launch_sweep_k<DP, BO_KERNEL_RBF, true>(h, a, grid, st) : launch_sweep_k<DP, BO_KERNEL_RBF, false>(h, a, grid, st);
    }
}
template <int DP>
static int launch_sweep_ref(bo_handle* h, const SweepArgs& a, cudaStream_t st) {
    const size_t sm = (size_t)a.np * sizeof(double);
    BO_CUDA(h, cudaFuncSetAttribute(sweep_reference_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    sweep_reference_kernel<DP><<<(unsigned)a.N, 256, sm, st>>>(a, h->Li, h->cap_np);
    BO_LAUNCH_CHECK(h);
    return 0;
}
template <int DP>
static int launch_sobol_points(bo_handle* h, const int64_t* idx, int64_t N, double* out, cudaStream_t st) {
    sobol_points_kernel<DP><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(h->sobol_dev, (const long long*)idx, N, h->d, out);
    BO_LAUNCH_CHECK(h);
    return 0;
}

#define BO_DISPATCH_DP(dp, fn, ...)                                   \
    ((dp) == 2 ? fn<2>(__VA_ARGS__) : (dp) == 4 ? fn<4>(__VA_ARGS__) :  \
     (dp) == 6 ? fn<6>(__VA_ARGS__) : (dp) == 8 ? fn<8>(__VA_ARGS__) :  \
     (dp) == 12 ? fn<12>(__VA_ARGS__) : fn<16>(__VA_ARGS__))

// The FP64 DMMA sweep of a prepared argument block (hyper-parameters, pool, outputs).  The per-CTA top-k lists go to
// slots [part_off, part_off + grid) of the handle's list buffers; `merge` reduces them to vals_dev / idx_dev.  Returns the
// grid through *grid_out.
static int sweep_fp64_run(bo_handle* h, SweepArgs a, int part_off, bool merge, double* vals_dev, int64_t* idx_dev, int* grid_out,
                          cudaStream_t st) {
    int rc;
    a.nblocks = (a.N + SW_BN - 1) / SW_BN;
    {
        const char* gs = getenv("BO_B200_SWEEP_SEGMENTS");      // test hook: force a segment count
        a.G = choose_segments(h->sm_count, a.nblocks, h->np / SW_BM, a.seg);
        if (h->svgp) {                 // the second pass needs the whole interp term of a block in one CTA: no row split
            a.G = 1; a.seg[0] = 0; a.seg[1] = h->np / SW_BM;
        } else if (gs && atoi(gs) >= 1) {
            const int nbm = h->np / SW_BM;
            int G = atoi(gs); if (G > nbm) G = nbm; if (G > SW_MAX_SEG) G = SW_MAX_SEG;
            // re-use the balancing by pretending a pool that makes G optimal: simple equal-stage split
            const double total = 0.5 * nbm * (nbm + 1.0);
            a.seg[0] = 0;
            for (int s2 = 1; s2 < G; ++s2) {
                int b = (int)floor((-1.0 + sqrt(1.0 + 8.0 * total * s2 / G)) / 2.0 + 0.5);
                if (b <= a.seg[s2 - 1]) b = a.seg[s2 - 1] + 1;
                if (b > nbm - (G - s2)) b = nbm - (G - s2);
                a.seg[s2] = b;
            }
            a.seg[G] = nbm; a.G = G;
        }
    }
    const long long items = a.nblocks * a.G;
    const int grid = (int)(items < h->sm_count ? items : h->sm_count);
    if ((rc = ensure_sweep_ws(h, grid, true, part_off + grid))) return rc;
    if (a.G > 1) {
        const int nbm = h->np / SW_BM;
        if ((rc = ensure_split_ws(h, a.nblocks, a.G, nbm))) return rc;
        a.part_cs = reinterpret_cast<double*>(h->split_ws);              // [nblocks][nbm][2][128] row-block sums
        a.part_mu = a.part_cs + (size_t)a.nblocks * nbm * 2 * SW_BN;
        a.part_kss = a.part_mu + (size_t)a.nblocks * SW_BN;
        a.counters = reinterpret_cast<int*>(a.part_kss + (size_t)a.nblocks * SW_BN);
        BO_CUDA(h, cudaMemsetAsync(a.counters, 0, (size_t)a.nblocks * sizeof(int), st));
    }
    a.panel = h->panel;
    a.part_val = h->part_val + (size_t)part_off * BO_MAX_TOPK; a.part_idx = (long long*)h->part_idx + (size_t)part_off * BO_MAX_TOPK;
    if (h->svgp) { a.Lp2 = h->Lp2; a.panel2 = h->panel2; a.sv_add = h->sv_add; }
    if (merge) BO_CUDA(h, cudaEventRecord(h->ev0, st));
    if ((rc = BO_DISPATCH_DP(h->dp, launch_sweep, h, a, grid, st))) return rc;
    if (merge) {
        BO_CUDA(h, cudaEventRecord(h->ev1, st));
        h->sweep_timed = true; h->sweep_path = 0; h->sweep_flagged = 0;
        if (a.topk > 0) {
            topk_merge_kernel<<<1, 1024, 0, st>>>(h->part_val, (long long*)h->part_idx, (part_off + grid) * BO_MAX_TOPK, a.topk, vals_dev, (long long*)idx_dev);
            BO_LAUNCH_CHECK(h);
        }
    }
    if (grid_out) *grid_out = grid;
    return 0;
}

#include "sweep_i8.cuh"

static int upload_sobol(bo_handle* h, const bo_sobol* sobol_host, cudaStream_t st) {
    if (sobol_host->d < h->d) return fail(h, BO_E_INVALID, "sobol state has fewer dimensions than the fitted model");
    if (!h->sobol_dev) BO_CUDA(h, cudaMalloc(&h->sobol_dev, sizeof(bo_sobol)));
    BO_CUDA(h, cudaMemcpyAsync(h->sobol_dev, sobol_host, sizeof(bo_sobol), cudaMemcpyHostToDevice, st));
    BO_CUDA(h, cudaStreamSynchronize(st));     // caller's struct is pageable and borrowed for the call only
    return 0;
}

int sweep_impl(bo_handle* h, int acq_kind, double best_f, double beta, double min_var,
               const double* cand_dev, const bo_sobol* sobol_host, int64_t first_index, int64_t N,
               int topk, double* vals_dev, int64_t* idx_dev, double* mean_dev, double* var_dev,
               double* acq_dev, cudaStream_t st, int mode_override) {
    if (!h->fitted) return fail(h, BO_E_NOTFIT, "sweep before a successful bo_fit");
    if (N < 0 || first_index < 0 || topk < 0) return fail(h, BO_E_INVALID, "bo_sweep: negative size");
    if (topk > BO_MAX_TOPK) return fail(h, BO_E_CAPACITY, "bo_sweep: topk exceeds BO_MAX_TOPK");
    if (acq_kind < BO_ACQ_EI || acq_kind > BO_ACQ_MEAN) return fail(h, BO_E_INVALID, "bo_sweep: unknown acquisition kind");
    if (!cand_dev && !sobol_host && N > 0) return fail(h, BO_E_INVALID, "bo_sweep: neither candidates nor a Sobol state given");
    if (topk > 0 && (!vals_dev || !idx_dev)) return fail(h, BO_E_INVALID, "bo_sweep: topk outputs missing");
    if (!(beta >= 0.0)) return fail(h, BO_E_INVALID, "bo_sweep: beta must be >= 0");
    if (!cand_dev && N > 0 && first_index + N > (1LL << BO_SOBOL_BITS))
        return fail(h, BO_E_CAPACITY, "bo_sweep: the 30-bit Sobol pool holds 2^30 points (torch.quasirandom.SobolEngine.MAXBIT)");
    BO_CUDA(h, cudaSetDevice(h->device));
    int rc;
    if (!cand_dev && N > 0 && (rc = upload_sobol(h, sobol_host, st))) return rc;

    SweepArgs a{};
    a.Xs = h->Xs; a.alpha = h->alpha; a.Lp = h->Lp;
    a.cand = cand_dev; a.sobol = cand_dev ? nullptr : h->sobol_dev;
    a.first_index = first_index; a.N = N; a.nblocks = (N + SW_BN - 1) / SW_BN;
    a.n = h->n; a.np = h->np; a.d = h->d; a.hyp = h->hyp;
    a.acq = acq_kind; a.best_f = best_f; a.sqrt_beta = sqrt(beta); a.min_var = min_var;
    a.topk = topk; a.mean_out = mean_dev; a.var_out = var_dev; a.acq_out = acq_dev;
    // bit 1: skip structurally-zero diagonal tiles (FP64 kernel); bit 2: warp-role cycle accounting (sliced kernels, triage);
    // bits 3-6: L2 eviction priority of the CTA-pair kernel's TMA loads -- default 8 | 64: the packed L^-1 tiles (read by every CTA pair
    // of the grid) evict_last, the per-pair panel tiles evict_first: +2.6 % candidates/s at C3 on a power-bound kernel (fewer DRAM
    // re-reads of L^-1, SM clock 1492 -> 1515 MHz under the same cap; each hint alone, or the opposite pairing, measured no gain:
    // profiles/r02_l2_hint_ab.log)
    { const char* f = getenv("BO_B200_SWEEP_FLAGS"); a.flags = f ? atoi(f) : (2 | 8 | 64); }

    if (N == 0) {
        if (topk > 0) {
            if ((rc = ensure_sweep_ws(h, 1))) return rc;
            score_topk_kernel<<<1, 256, 0, st>>>(nullptr, 0, first_index, topk, h->part_val, (long long*)h->part_idx);
            BO_LAUNCH_CHECK(h);
            topk_merge_kernel<<<1, 1024, 0, st>>>(h->part_val, (long long*)h->part_idx, BO_MAX_TOPK, topk, vals_dev, (long long*)idx_dev);
            BO_LAUNCH_CHECK(h);
        }
        return 0;
    }

    const char* impl = getenv("BO_B200_SWEEP_IMPL");
    if (impl && strcmp(impl, "reference") == 0 && !h->svgp) {
        // slow independent path: needs dense outputs; borrow temporaries if the caller passed none
        double *tm = nullptr, *tv = nullptr, *ta = nullptr;
        if (!a.mean_out) { BO_CUDA(h, cudaMalloc(&tm, N * 8)); a.mean_out = tm; }
        if (!a.var_out)  { BO_CUDA(h, cudaMalloc(&tv, N * 8)); a.var_out = tv; }
        if (!a.acq_out)  { BO_CUDA(h, cudaMalloc(&ta, N * 8)); a.acq_out = ta; }
        const int grid = (int)(N < 64 ? N : 64);
        if ((rc = ensure_sweep_ws(h, grid))) return rc;
        if ((rc = BO_DISPATCH_DP(h->dp, launch_sweep_ref, h, a, st))) return rc;
        if (topk > 0) {
            score_topk_kernel<<<grid, 256, 0, st>>>(a.acq_out, N, first_index, topk, h->part_val, (long long*)h->part_idx);
            BO_LAUNCH_CHECK(h);
            topk_merge_kernel<<<1, 1024, 0, st>>>(h->part_val, (long long*)h->part_idx, grid * BO_MAX_TOPK, topk, vals_dev, (long long*)idx_dev);
            BO_LAUNCH_CHECK(h);
        }
        BO_CUDA(h, cudaStreamSynchronize(st));
        if (tm) cudaFree(tm); if (tv) cudaFree(tv); if (ta) cudaFree(ta);
        return 0;
    }

    {
        // which contraction: the handle's mode (bo_set_sweep_mode), overridable for triage by BO_B200_SWEEP_IMPL=fp64|i8
        int mode = mode_override >= 0 ? mode_override : h->sweep_mode;
        if (impl && strcmp(impl, "fp64") == 0) mode = BO_SWEEP_FP64;
        if (impl && strcmp(impl, "i8") == 0) {
            const char* sl = getenv("BO_B200_I8_SLICES");
            mode = (sl && atoi(sl) == 8) ? BO_SWEEP_I8X8 : (sl && atoi(sl) == 7) ? BO_SWEEP_I8X7 : BO_SWEEP_AUTO;
        }
        mode = resolve_sweep_mode(h, mode, N);
        if (mode != BO_SWEEP_FP64) {
            const int S = mode == BO_SWEEP_I8X7 ? 7 : 8;
            return sweep_i8_run(h, a, S, vals_dev, idx_dev, st);
        }
    }

    return sweep_fp64_run(h, a, 0, true, vals_dev, idx_dev, nullptr, st);
}

int sobol_points_impl(bo_handle* h, const bo_sobol* sobol_host, const int64_t* idx_dev, int64_t N,
                      double* out_dev, cudaStream_t st) {
    if (!sobol_host || !idx_dev || !out_dev || N < 0) return fail(h, BO_E_INVALID, "bo_sobol_points: bad argument");
    if (h->d <= 0) return fail(h, BO_E_NOTFIT, "bo_sobol_points: dimension unknown before bo_fit");
    if (N == 0) return 0;
    BO_CUDA(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = upload_sobol(h, sobol_host, st))) return rc;
    return BO_DISPATCH_DP(pad_dim(h->d), launch_sobol_points, h, idx_dev, N, out_dev, st);
}

}  // namespace bo
