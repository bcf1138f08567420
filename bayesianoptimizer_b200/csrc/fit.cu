// GP fit on device: fused pairwise-distance + Matern-5/2/RBF Gram builder,
// FP64 blocked right-looking Cholesky (DMMA SYRK/GEMM trailing updates), explicit L^-1 by
// recursive block inversion, alpha with one step of iterative refinement, and the repack of
// L^-1 into the sweep kernel's DMMA fragment order.
// Replaces SingleTaskGP construction + prediction caches (optimization/Bayesian.py:89-94).
#include "gemm.cuh"

namespace bo {

int launch_trmv_lower(bo_handle* h, const double* v, double* z, cudaStream_t st);
int launch_trmv_lower_t(bo_handle* h, const double* z, double* out, int accumulate, cudaStream_t st);
static int trmv_lower_m(bo_handle* h, const double* M, const double* v, double* z, cudaStream_t st);
static int trmv_lower_t_m(bo_handle* h, const double* M, const double* z, double* out, int accumulate, cudaStream_t st);

// ------------------------------------------------------------------------------------------
// input staging: Xraw (unscaled, padded to BO_MAX_DIM columns), Xs = Xraw * inv_ls, y
// ------------------------------------------------------------------------------------------
__global__ void stage_inputs_kernel(const double* __restrict__ X, const double* __restrict__ y, int n, int d,
                                    int np, Hyper hyp, double* __restrict__ Xraw, double* __restrict__ Xs,
                                    double* __restrict__ yv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
#pragma unroll
    for (int k = 0; k < BO_MAX_DIM; ++k) {
        double v = (i < n && k < d) ? X[(size_t)i * d + k] : 0.0;
        Xraw[(size_t)i * BO_MAX_DIM + k] = v;
        Xs[(size_t)i * BO_MAX_DIM + k] = (k < d) ? v * hyp.inv_ls[k] : 0.0;
    }
    yv[i] = (i < n) ? y[i] : 0.0;
}

// ------------------------------------------------------------------------------------------
// K1: Gram builder, lower triangle of K = k(X,X) + (noise + jitter) I, identity on the padding
// ------------------------------------------------------------------------------------------
// CTA tile: GRAM_ROWS rows x 32 columns (lane = column, warp = 16 consecutive rows, processed 4 at a time with
// straight-line code so that the sqrt/exp chains of independent pairs interleave); the kernel kind is resolved once
// outside the loops.  Only tiles that touch the lower triangle do work.
constexpr int GRAM_ROWS = 128;

template <int DP, int KIND>
__device__ __forceinline__ void gram_rows(const double (*xs_i)[DP], const double* xj, int i0, int j, int n, int ld,
                                          const Hyper& hyp, double* __restrict__ K, int warp) {
    double xw[KIND == BO_KERNEL_LINEAR_MATERN52 ? DP : 1];
    if (KIND == BO_KERNEL_LINEAR_MATERN52) {
#pragma unroll
        for (int k = 0; k < DP; ++k) xw[k] = hyp.lin_w[k] * xj[k];
    }
    for (int rb = 0; rb < 16; rb += 4) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = warp * 16 + rb + u;
            double sq = 0.0, lin = 0.0;
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                const double xi = xs_i[r][k];
                const double df = xi - xj[k];
                sq = fma(df, df, sq);
                if (KIND == BO_KERNEL_LINEAR_MATERN52) lin = fma(xw[k], xi, lin);
            }
            const int i = i0 + r;
            const double off = kernel_pair_t<KIND>(sq, lin, hyp.outputscale);
            const double dg = hyp.outputscale * (KIND == BO_KERNEL_LINEAR_MATERN52 ? lin + 1.0 : 1.0) + hyp.noise + hyp.jitter;   // exact diagonal
            const double inside = (i == j) ? dg : off;
            v[u] = (i >= n || j >= n) ? ((i == j) ? 1.0 : 0.0) : inside;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + warp * 16 + rb + u;
            if (j <= i) K[(size_t)i * ld + j] = v[u];
        }
    }
}

template <int DP>
__device__ __forceinline__ void gram_body(const double* __restrict__ Xs, int n, int np, int ld,
                                          const Hyper& hyp, double* __restrict__ K) {
    // block = 32 (cols) x 8 (warps) threads, tile GRAM_ROWS x 32
    __shared__ double xs_i[GRAM_ROWS][DP];
    const int bj = blockIdx.x, bi = blockIdx.y;
    if (bj * 32 > bi * GRAM_ROWS + GRAM_ROWS - 1) return;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int i0 = bi * GRAM_ROWS, j = bj * 32 + threadIdx.x;
    for (int e = tid; e < GRAM_ROWS * DP; e += 256) xs_i[e / DP][e % DP] = Xs[(size_t)(i0 + e / DP) * BO_MAX_DIM + e % DP];
    double xj[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) xj[k] = Xs[(size_t)j * BO_MAX_DIM + k];
    __syncthreads();
    if (hyp.kind == BO_KERNEL_MATERN52) gram_rows<DP, BO_KERNEL_MATERN52>(xs_i, xj, i0, j, n, ld, hyp, K, threadIdx.y);
    else if (hyp.kind == BO_KERNEL_RBF) gram_rows<DP, BO_KERNEL_RBF>(xs_i, xj, i0, j, n, ld, hyp, K, threadIdx.y);
    else gram_rows<DP, BO_KERNEL_LINEAR_MATERN52>(xs_i, xj, i0, j, n, ld, hyp, K, threadIdx.y);
}

template <int DP>
__global__ void __launch_bounds__(256, 3) gram_kernel(const double* __restrict__ Xs, int n, int np, int ld,
                                                   Hyper hyp, double* __restrict__ K) {
    gram_body<DP>(Xs, n, np, ld, hyp, K);
}
// per-slot variant for the batched LML restarts: blockIdx.z = slot, hyper-parameters from a device array
template <int DP>
__global__ void __launch_bounds__(256, 3) gram_batched_kernel(const double* __restrict__ Xs, int n, int np, int ld,
                                                           const Hyper* __restrict__ hyps, double* __restrict__ K) {
    const size_t s = blockIdx.z;
    gram_body<DP>(Xs + s * np * BO_MAX_DIM, n, np, ld, hyps[s], K + s * np * ld);
}
__global__ void rescale_batched_kernel(int np, int d, const Hyper* __restrict__ hyps, const double* __restrict__ Xraw,
                                       double* __restrict__ Xs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const Hyper& hy = hyps[blockIdx.z];
    double* dst = Xs + ((size_t)blockIdx.z * np + i) * BO_MAX_DIM;
#pragma unroll
    for (int k = 0; k < BO_MAX_DIM; ++k) dst[k] = (k < d) ? Xraw[(size_t)i * BO_MAX_DIM + k] * hy.inv_ls[k] : 0.0;
}

// ------------------------------------------------------------------------------------------
// K2 panel: Cholesky of the NB x NB diagonal block fused with the triangular solve of the rows below it
// ------------------------------------------------------------------------------------------
// Right-looking Cholesky with the block held in REGISTERS: thread (ti, tk) of a 16 x 16 grid owns the
// 4 x 4 elements (ti + 16x, tk + 16y); per column one shared-memory broadcast of the pivot column and one
// barrier (double-buffered), the rank-1 update is 16 register FMAs.  Leaves L (lower, upper zero) in S.
// The update is UNMASKED: a finished column is copied to S the moment it is final, after which the registers of
// rows <= j / columns <= j may hold anything -- a valid entry (i, k > j) is only ever updated with L[i][j] L[k][j],
// both taken from the still-valid part of the pivot column, so stale values never reach a live entry.  This halves
// the instructions of a step (the chain of 64 steps is issue-bound: 124 -> ~65 SASS instructions per warp and step).
// BELOW = true: the CTA also carries 64 rows P of the panel below the diagonal block through the SAME rank-1 updates (the
// right-looking form of X L^T = P: once column j of L is final, x_j = p_j / L[j][j] and p_k -= x_j L[k][j] for k > j), so the
// triangular solve costs 16 more FMAs per step instead of a second chain of 64 dependent steps; S then receives X, not L.
template <bool BELOW>
__device__ __forceinline__ void leaf_cholesky(const double* __restrict__ A, int lda, double (*S)[NB + 1],
                                              double (*colS)[NB], int* info, int pivot_base, bool report,
                                              const double* __restrict__ P = nullptr, double (*colB)[NB] = nullptr, int kpre = 0) {
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tk = tid & 15;
    double a[4][4];
    double bw[BELOW ? 4 : 1][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int i = ti + 16 * x, k = tk + 16 * y;
            a[x][y] = (k <= i) ? A[(size_t)i * lda + k] : 0.0;
            if (BELOW) bw[x][y] = P[(size_t)i * lda + k];
        }
    // LEFT-LOOKING inside an outer block (kpre > 0): the kpre columns to the left of this block column (the earlier panels of the
    // outer block, final) have not been applied to it -- no narrow K = 64 GEMM launch between two panels -- so this CTA subtracts
    // their product from its copy of the diagonal block and from its own rows first: D -= Dl Dl^T, P -= Pl Dl^T with
    // Dl = A[:, -kpre:0), Pl = P[:, -kpre:0).  Chunks of 32 columns are staged k-major in S (unused until a column is final).
    if (kpre > 0) {
        double (*T0)[NB + 1] = S;                       // [32][65]: Dl chunk, T0[k][row]
        double (*T1)[NB + 1] = S + 32;                  // [32][65]: Pl chunk
        for (int k0 = -kpre; k0 < 0; k0 += 32) {
            __syncthreads();
            for (int e = tid; e < NB * 32; e += 256) {
                const int row = e >> 5, k = e & 31;
                const ptrdiff_t off = (ptrdiff_t)row * lda + k0 + k;        // k0 < 0: columns left of the block
                T0[k][row] = A[off];
                if (BELOW) T1[k][row] = P[off];
            }
            __syncthreads();
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                double dI[4], dK[4], pI[BELOW ? 4 : 1];
#pragma unroll
                for (int x = 0; x < 4; ++x) { dI[x] = T0[k][ti + 16 * x]; if (BELOW) pI[x] = T1[k][ti + 16 * x]; }
#pragma unroll
                for (int y = 0; y < 4; ++y) dK[y] = T0[k][tk + 16 * y];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        a[x][y] = fma(-dI[x], dK[y], a[x][y]);
                        if (BELOW) bw[x][y] = fma(-pI[x], dK[y], bw[x][y]);
                    }
            }
        }
        __syncthreads();                                 // S is free again before the first finished column lands in it
    }
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {            // unrolled: every register index below is static
#pragma unroll 1
        for (int jt = 0; jt < 16; ++jt) {
            const int j = jb * 16 + jt;
            double* col = colS[j & 1];
            if (tk == jt) {
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    col[ti + 16 * x] = a[x][jb];
                    if (BELOW) colB[j & 1][ti + 16 * x] = bw[x][jb];
                }
            }
            __syncthreads();
            double dj = col[j];
            if (!(dj > 0.0)) {                       // also catches NaN
                if (report && tid == 0) atomicCAS(info, 0, pivot_base + j + 1);
                dj = 1.0;
            }
            double rs;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs) : "d"(dj));
            const double hj = 0.5 * dj;
            rs = rs * fma(-hj, rs * rs, 1.5);
            rs = rs * fma(-hj, rs * rs, 1.5);
            double li[4], lk[4], lb[BELOW ? 4 : 1];
#pragma unroll
            for (int x = 0; x < 4; ++x) li[x] = col[ti + 16 * x] * rs;
#pragma unroll
            for (int y = 0; y < 4; ++y) lk[y] = col[tk + 16 * y] * rs;
            if (BELOW) {
#pragma unroll
                for (int x = 0; x < 4; ++x) lb[x] = colB[j & 1][ti + 16 * x] * rs;
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
                    if (y >= jb) {                                           // column groups left of j are final
                        a[x][y] = fma(-li[x], lk[y], a[x][y]);
                        if (BELOW) bw[x][y] = fma(-lb[x], lk[y], bw[x][y]);
                    }
            if (BELOW) {
                if (tk == jt) {                      // column j of X is final
#pragma unroll
                    for (int x = 0; x < 4; ++x) S[ti + 16 * x][j] = lb[x];
                }
            } else
            if (tk == jt) {                          // column j is final: park it in S (zero above the diagonal)
                double sj = dj * rs;
                sj = fma(0.5 * rs, fma(-sj, sj, dj), sj);          // sqrt(dj), Heron-corrected
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int i = ti + 16 * x;
                    S[i][j] = (i > j) ? li[x] : (i == j ? sj : 0.0);
                }
            }
        }
    }
    __syncthreads();
}

// grid = slots * ctas_per_slot; CTA r of a slot: r == 0 factors the diagonal block and parks L in the (still
// unused) diagonal block of the inverse matrix -- NOT in place, later-scheduled CTAs of this launch still read
// the unfactored block -- and r >= 1 re-factors it (identical arithmetic, no inter-CTA wait) and solves rows
// [64 r, 64 r + 64) of the panel below: X L^T = A by forward substitution, 4 threads per row.
// Replaces the potrf2 + TRSM launches of a block column.
__global__ void __launch_bounds__(256) chol_panel_kernel(double* __restrict__ A, int lda, double* __restrict__ Lpark,
                                                         int* __restrict__ info, int pivot_base, size_t slot_stride,
                                                         int ctas_per_slot, int fused, int kpre) {
    __shared__ double S[NB][NB + 1];
    __shared__ double colS[2][NB];
    __shared__ double colB[2][NB];
    __shared__ double rdiag[NB];
    const int slot = blockIdx.x / ctas_per_slot, r = blockIdx.x % ctas_per_slot;
    A += slot * slot_stride; Lpark += slot * slot_stride; info += slot;
    const int tid = threadIdx.x;
    if (r == 0) {
        leaf_cholesky<false>(A, lda, S, colS, info, pivot_base, true, nullptr, nullptr, kpre);
        for (int e = tid; e < NB * NB; e += 256) {
            const int i = e / NB, k = e % NB;
            if (k <= i) Lpark[(size_t)i * lda + k] = S[i][k];
        }
        return;
    }
    if (fused) {
        // the rows below ride through the factorisation's own rank-1 updates: one chain of 64 steps instead of two
        double* Pf = A + (size_t)r * NB * lda;
        leaf_cholesky<true>(A, lda, S, colS, info, pivot_base, false, Pf, colB, kpre);
        for (int e = tid; e < NB * NB; e += 256) {
            const int i = e / NB, k = e % NB;
            Pf[(size_t)i * lda + k] = S[i][k];
        }
        return;
    }
    leaf_cholesky<false>(A, lda, S, colS, info, pivot_base, false);
    if (tid < NB) rdiag[tid] = 1.0 / S[tid][tid];
    __syncthreads();
    double* P = A + (size_t)r * NB * lda;                  // 64 rows of the panel below the diagonal block
    const int row = tid >> 2, q = tid & 3, lane = tid & 31;
    double av[NB / 4], x[NB / 4];
#pragma unroll
    for (int m = 0; m < NB / 4; ++m) { av[m] = P[(size_t)row * lda + 4 * m + q]; x[m] = 0.0; }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        double pp[4] = {0.0, 0.0, 0.0, 0.0};                  // four independent chains: FP64 latency, not issue, binds here
#pragma unroll
        for (int m = 0; m < NB / 4; ++m)
            if (4 * m + q < j) pp[m & 3] = fma(x[m], S[j][4 * m + q], pp[m & 3]);
        double part = (pp[0] + pp[1]) + (pp[2] + pp[3]);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        const double aj = __shfl_sync(0xffffffffu, av[j >> 2], (lane & ~3) | (j & 3));
        const double xj = (aj - part) * rdiag[j];
        if ((j & 3) == q) x[j >> 2] = xj;
    }
#pragma unroll
    for (int m = 0; m < NB / 4; ++m) P[(size_t)row * lda + 4 * m + q] = x[m];
}

// explicit inverses of all NB x NB diagonal blocks of L in one launch (grid = slots * nb), off the critical path.
// Reads the factored block parked in Linv's diagonal block, moves it to its place in L, and overwrites the
// parking spot with the inverse: forward substitution, 4 threads per column c, thread q owns rows i % 4 == q.
__global__ void __launch_bounds__(256) leaf_inverse_kernel(double* __restrict__ L, int ld, double* __restrict__ Linv,
                                                           size_t slot_stride, int nb) {
    __shared__ double S[NB][NB + 1];
    __shared__ double rdiag[NB];
    const int slot = blockIdx.x / nb, kb = blockIdx.x % nb;
    const size_t off = slot * slot_stride + (size_t)kb * NB * ld + kb * NB;
    double* A = L + off;
    double* Ainv = Linv + off;
    const int tid = threadIdx.x;
    for (int e = tid; e < NB * NB; e += 256) {
        const int i = e / NB, k = e % NB;
        const double v = (k <= i) ? Ainv[(size_t)i * ld + k] : 0.0;
        S[i][k] = v;
        if (k <= i) A[(size_t)i * ld + k] = v;
    }
    __syncthreads();
    if (tid < NB) rdiag[tid] = 1.0 / S[tid][tid];
    __syncthreads();
    const int c = tid >> 2, q = tid & 3;
    double x[NB / 4];
#pragma unroll
    for (int m = 0; m < NB / 4; ++m) x[m] = 0.0;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        double pp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int m = 0; m < NB / 4; ++m) {
            if (4 * m + q < i) pp[m & 3] = fma(S[i][4 * m + q], x[m], pp[m & 3]);
        }
        double part = (pp[0] + pp[1]) + (pp[2] + pp[3]);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        const double xi = ((i == c ? 1.0 : 0.0) - part) * rdiag[i];
        if ((i & 3) == q) x[i >> 2] = xi;
    }
#pragma unroll
    for (int m = 0; m < NB / 4; ++m) Ainv[(size_t)(4 * m + q) * ld + c] = x[m];
}

// ------------------------------------------------------------------------------------------
// K3 vectors: triangular mat-vecs with the explicit inverse / the factor (alpha + its refinement)
// ------------------------------------------------------------------------------------------
__global__ void resid_init_kernel(const double* __restrict__ y, int n, int np, double mean, double* __restrict__ r) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    r += (size_t)blockIdx.z * np;                       // slot (y is shared by all slots)
    if (i < np) r[i] = (i < n) ? y[i] - mean : 0.0;
}

// z[i] = sum_{j<=i} Li[i][j] v[j]   (one warp per row)
__global__ void __launch_bounds__(256) trmv_lower_kernel(const double* __restrict__ Li, int ld, int np,
                                                         const double* __restrict__ v, double* __restrict__ z,
                                                         size_t mat_stride) {
    Li += blockIdx.z * mat_stride; v += (size_t)blockIdx.z * np; z += (size_t)blockIdx.z * np;      // slot
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= np) return;
    // 16-byte loads, four independent pairs of accumulators: 2 KB of the row in flight per warp (the first version issued one
    // 8-byte load per lane and iteration: 36 us for the 71 MB of an n = 4224 factor, 2 TB/s; rows and vectors are 16-byte aligned)
    const double2* a2 = reinterpret_cast<const double2*>(Li + (size_t)row * ld);
    const double2* v2 = reinterpret_cast<const double2*>(v);
    const int npair = (row + 1) >> 1;                   // pairs (2p, 2p + 1) entirely inside j <= row
    double sx[4] = {0.0, 0.0, 0.0, 0.0}, sy[4] = {0.0, 0.0, 0.0, 0.0};
    int p = lane;
    for (; p + 96 < npair; p += 128) {
        double2 av[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { av[u] = a2[p + 32 * u]; vv[u] = v2[p + 32 * u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { sx[u] = fma(av[u].x, vv[u].x, sx[u]); sy[u] = fma(av[u].y, vv[u].y, sy[u]); }
    }
    for (; p < npair; p += 32) { const double2 av = a2[p], vv = v2[p]; sx[0] = fma(av.x, vv.x, sx[0]); sy[0] = fma(av.y, vv.y, sy[0]); }
    double s = ((sx[0] + sx[1]) + (sx[2] + sx[3])) + ((sy[0] + sy[1]) + (sy[2] + sy[3]));
    if (lane == 0 && !(row & 1)) s = fma(Li[(size_t)row * ld + row], v[row], s);       // odd count: the diagonal element is left over
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) z[row] = s;
}

// out[j] (+)= sum_{i>=j} Li[i][j] z[i]: block (x = 32 columns, y = row split), partial sums per split, then a
// deterministic reduction over the splits
constexpr int TRMVT_SPLITS = 16;
__global__ void __launch_bounds__(256) trmv_lower_t_kernel(const double* __restrict__ Li, int ld, int np,
                                                           const double* __restrict__ z, double* __restrict__ part,
                                                           size_t mat_stride) {
    Li += blockIdx.z * mat_stride; z += (size_t)blockIdx.z * np; part += (size_t)blockIdx.z * TRMVT_SPLITS * np;   // slot
    __shared__ double red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j0 = blockIdx.x * 32, j = j0 + tx;
    // rows j0 .. np split into TRMVT_SPLITS contiguous chunks (multiples of 8)
    const int rows = np - j0;
    const int chunk = ((rows + TRMVT_SPLITS - 1) / TRMVT_SPLITS + 7) & ~7;
    const int r0 = j0 + blockIdx.y * chunk, r1 = min(np, r0 + chunk);
    double s = 0.0;
    for (int i = r0 + ty; i < r1; i += 8)          // (an explicit four-rows-in-flight form of this loop measured 50 % slower)
        if (i >= j) s = fma(Li[(size_t)i * ld + j], z[i], s);
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][tx];
        part[(size_t)blockIdx.y * np + j] = t;
    }
}
__global__ void trmv_reduce_kernel(const double* __restrict__ part, int np, double* __restrict__ out, int accumulate) {
    part += (size_t)blockIdx.z * TRMVT_SPLITS * np; out += (size_t)blockIdx.z * np;                    // slot
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= np) return;
    double t = 0.0;
#pragma unroll
    for (int s = 0; s < TRMVT_SPLITS; ++s) t += part[(size_t)s * np + j];
    out[j] = accumulate ? out[j] + t : t;
}

// ------------------------------------------------------------------------------------------
// repack L^-1 (row-major) into the sweep's A-operand stage tiles:
//   tile (ib, kc), kc < (ib+1)*SW_BM/SW_BK, at index (ib*(ib+1)/2)*(SW_BM/SW_BK) + kc, SW_TILE doubles each;
//   inside a tile the 8x8 sub-block (r8, k8) sits at (r8*(SW_BK/8) + k8)*64 and its element (r, k)
//   at lane*2 + khalf with lane = (r%8)*4 + k%4, khalf = (k%8)/4  -> one LDS.128 per lane fetches the
//   DMMA.8x8x4 A fragments of two consecutive k4 steps.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_linv_kernel(const double* __restrict__ Li, int ld, int np,
                                                        double* __restrict__ Lp, int ib0) {
    constexpr int KCH = SW_BM / SW_BK;
    const int ib = ib0 + blockIdx.y;
    const int kc = blockIdx.x;
    if (kc >= (ib + 1) * KCH) return;
    double* dst = Lp + ((size_t)ib * (ib + 1) / 2 * KCH + kc) * SW_TILE;
    const int row0 = ib * SW_BM, k0 = kc * SW_BK;
    for (int e = threadIdx.x; e < SW_TILE; e += 256) {
        const int r = e / SW_BK, k = e % SW_BK;          // coalesced 256 B reads along k
        const double v = (k0 + k <= row0 + r) ? Li[(size_t)(row0 + r) * ld + k0 + k] : 0.0;
        const int r8 = r >> 3, k8 = k >> 3;
        const int lane = (r & 7) * 4 + (k & 3), khalf = (k & 7) >> 2;
        dst[(r8 * (SW_BK / 8) + k8) * 64 + lane * 2 + khalf] = v;
    }
}

// copy lower triangle (upper zero) of the leading n x n block to a dense n x n output
__global__ void export_lower_kernel(const double* __restrict__ src, int ld, int n, double* __restrict__ dst) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    int i = (int)(e / n), j = (int)(e % n);
    dst[e] = (j <= i) ? src[(size_t)i * ld + j] : 0.0;
}

// ------------------------------------------------------------------------------------------
// grouped GEMM launcher
// ------------------------------------------------------------------------------------------
int gemm_init(bo_handle* h) {
    BO_CUDA(h, cudaFuncSetAttribute(dgemm_grouped_kernel<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)GemmSmem<64, 64>::BYTES));
    // production shape: 64x64 tiles, 2-stage ring, 4 CTAs/SM (64 registers) -- best at every K in tools/gemm_probe.py
    BO_CUDA(h, cudaFuncSetAttribute(dgemm_grouped_kernel<64, 64, 16, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)GemmSmem<64, 64, 16, 2>::BYTES));
    BO_CUDA(h, cudaFuncSetAttribute(dgemm_grouped_kernel<128, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)GemmSmem<128, 128>::BYTES));
    BO_CUDA(h, cudaFuncSetAttribute(dgemm_grouped_kernel<128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)GemmSmem<128, 64>::BYTES));
    return 0;
}

int gemm_launch(bo_handle* h, const GemmLaunch& L, cudaStream_t st) {
    if (L.tiles == 0) return 0;
    if (L.cfg == 1)
        dgemm_grouped_kernel<128, 128><<<L.tiles, 256, GemmSmem<128, 128>::BYTES, st>>>(h->plan_dev + L.first, L.count);
    else
        dgemm_grouped_kernel<64, 64, 16, 2, 4><<<L.tiles, 256, GemmSmem<64, 64, 16, 2>::BYTES, st>>>(h->plan_dev + L.first, L.count);
    BO_LAUNCH_CHECK(h);
    return 0;
}

static void plan_push(bo_handle* h, const GemmBatch& b) {
    GemmLaunch L{(int)h->plan_probs.size(), (int)b.probs.size(), b.tiles, b.bm == 128 ? 1 : 0};
    h->plan_probs.insert(h->plan_probs.end(), b.probs.begin(), b.probs.end());
    h->plan_launches.push_back(L);
}

// pick the tile: 128 when every dimension allows it and the launch still fills the GPU
static int pick_tile(int sm, std::initializer_list<int> dims, long tiles128) {
    // measured with bo_gemm_probe on B200 (tools/gemm_probe.py): 64x64 tiles (4 CTAs/SM with the 2-stage ring hide the
    // prologue and the C read-modify-write) beat 128x128 and 128x64 at every K from 64 to 4096 -> always 64
    (void)sm; (void)dims; (void)tiles128;
    return 64;
}

// Trailing update after the panel of block column kb in the TWO-LEVEL right-looking scheme: block columns are grouped
// into outer blocks of OB; inside an outer block a panel only updates the remaining columns of that block (K = NB,
// narrow), and the last panel of the block applies all OB panels to the rest of the matrix at once (K = OB * NB).
// The K = 64 updates stream C through L2 once per 64 columns (4 flop/B, 20 TFLOP/s); K = 256 quarters that traffic.
// BO_B200_PANEL_FUSED=0: the two-chain panel (factor the diagonal block, then forward-substitute the rows below) for A/B runs
int chol_panel_fused() {
    static int f = [] { const char* e = getenv("BO_B200_PANEL_FUSED"); return e ? (atoi(e) != 0 ? 1 : 0) : 1; }();
    return f;
}
// BO_B200_CHOL_LEFT=1 (A/B only, OFF by default): left-looking inside the outer blocks -- no narrow K = 64 GEMM launch after a panel,
// the next panel kernels apply it themselves.  Measured slower (refit 4.13 -> 4.66 ms at n = 4096, n = 3000 LML 3.08 -> 3.51 ms, C5
// 103.7 -> 108.7 ms): every CTA of a panel launch repeats the update of the diagonal block, and 64 x 64 x 64m products on plain FP64
// FMAs from shared memory cost more than the 12-23 us grouped-DMMA launches they replace.
int chol_left_looking() {
    static int f = [] { const char* e = getenv("BO_B200_CHOL_LEFT"); return (e ? (atoi(e) != 0 ? 1 : 0) : 0) && chol_panel_fused(); }();
    return f;
}
static int chol_outer_blocks();
// columns of the outer block to the left of block column kb that its panel kernel applies itself (left-looking form)
int chol_panel_kpre(int kb) { return chol_left_looking() ? (kb % chol_outer_blocks()) * NB : 0; }
static int chol_outer_blocks() {
    static int ob = [] { const char* e = getenv("BO_B200_CHOL_OB"); const int v = e ? atoi(e) : 4; return v >= 1 && v <= 16 ? v : 4; }();
    return ob;
}
static void add_trailing_update(GemmBatch& g, double* Lm, int ld, int np, int kb, int nb) {
    const int OB = chol_outer_blocks();
    const int ob0 = (kb / OB) * OB, oend = std::min(ob0 + OB, nb);
    const int r0 = (kb + 1) * NB;
    double* C = Lm + (size_t)r0 * ld + r0;
    if (kb + 1 < oend) {
        if (chol_left_looking()) return;                             // the later panels of the outer block apply panel kb themselves
        double* P = Lm + (size_t)r0 * ld + kb * NB;                  // panel kb below its diagonal block
        g.add(P, ld, P, ld, C, ld, np - r0, oend * NB - r0, NB, -1.0, 1.0, /*transB=*/1, GEMM_LOWER_C);
    } else if (r0 < np) {
        double* P = Lm + (size_t)r0 * ld + ob0 * NB;                 // panels ob0 .. kb side by side
        g.add(P, ld, P, ld, C, ld, np - r0, np - r0, (kb + 1 - ob0) * NB, -1.0, 1.0, /*transB=*/1, GEMM_LOWER_C);
    }
}

// Build the launch plan of the factorisation (one SYRK launch per block column; the panel kernel does the
// diagonal block + triangular solve) followed by the recursive inverse (2 launches per level).
//   plan_launches[kb] = SYRK trailing update of block column kb      (kb = 0 .. nb-2)
//   then per inverse level (deepest first): T = C * Ainv ; Li[lower-left] = -Binv * T
static int build_plan(bo_handle* h, cudaStream_t st) {
    const int np = h->np, ld = h->cap_np, nb = np / NB;
    h->plan_probs.clear();
    h->plan_launches.clear();
    for (int kb = 0; kb + 1 < nb; ++kb) {
        const int m = np - (kb + 1) * NB;
        long t128 = (long)(m / 128) * (m / 128 + 1) / 2;
        GemmBatch syrk(pick_tile(h->sm_count, {m}, t128));
        add_trailing_update(syrk, h->Lm, ld, np, kb, nb);
        plan_push(h, syrk);
    }
    // recursive inverse: collect merge nodes by depth
    struct Node { int lo, mid, hi, depth; };
    std::vector<Node> nodes;
    std::vector<Node> stack;
    int maxdepth = 0;
    if (nb > 1) stack.push_back({0, nb / 2, nb, 0});
    while (!stack.empty()) {
        Node nd = stack.back(); stack.pop_back();
        nodes.push_back(nd);
        if (nd.depth > maxdepth) maxdepth = nd.depth;
        if (nd.mid - nd.lo > 1) stack.push_back({nd.lo, nd.lo + (nd.mid - nd.lo) / 2, nd.mid, nd.depth + 1});
        if (nd.hi - nd.mid > 1) stack.push_back({nd.mid, nd.mid + (nd.hi - nd.mid) / 2, nd.hi, nd.depth + 1});
    }
    for (int depth = maxdepth; depth >= 0; --depth) {
        bool all128 = true; long t128 = 0;
        for (const Node& nd : nodes) if (nd.depth == depth) {
            int p = (nd.mid - nd.lo) * NB, q = (nd.hi - nd.mid) * NB;
            if (p % 128 || q % 128 || (nd.lo * NB) % 128) all128 = false;
            t128 += (long)(p / 128) * (q / 128);
        }
        const int tile = pick_tile(h->sm_count, {all128 ? 128 : 64}, t128);
        GemmBatch g1(tile), g2(tile);
        size_t toff = 0;
        for (const Node& nd : nodes) if (nd.depth == depth) {
            const int lo = nd.lo * NB, mid = nd.mid * NB, p = (nd.mid - nd.lo) * NB, q = (nd.hi - nd.mid) * NB;
            const double* Cblk = h->Lm + (size_t)mid * ld + lo;          // q x p
            const double* Ainv = h->Li + (size_t)lo * ld + lo;           // p x p lower
            const double* Binv = h->Li + (size_t)mid * ld + mid;         // q x q lower
            double* T = h->Tw + toff;                                     // q x p, ld = p
            double* Out = h->Li + (size_t)mid * ld + lo;
            g1.add(Cblk, ld, Ainv, ld, T, p, q, p, p, 1.0, 0.0, 0, GEMM_B_LOWER_NN);
            g2.add(Binv, ld, T, p, Out, ld, q, p, q, -1.0, 0.0, 0, GEMM_A_LOWER);
            toff += (size_t)q * p;
        }
        plan_push(h, g1);
        plan_push(h, g2);
    }
    const size_t bytes = h->plan_probs.size() * sizeof(GemmProblem);
    if (bytes > h->plan_dev_cap) {
        if (h->plan_dev) cudaFree(h->plan_dev);
        h->plan_dev = nullptr; h->plan_dev_cap = 0;
        BO_CUDA(h, cudaMalloc(&h->plan_dev, bytes + 4096));
        h->plan_dev_cap = bytes + 4096;
    }
    if (bytes) BO_CUDA(h, cudaMemcpyAsync(h->plan_dev, h->plan_probs.data(), bytes, cudaMemcpyHostToDevice, st));
    BO_CUDA(h, cudaStreamSynchronize(st));    // plan_probs is pageable host memory
    h->plan_np = np;
    return 0;
}

// ------------------------------------------------------------------------------------------
// capacity management
// ------------------------------------------------------------------------------------------
int ensure_capacity(bo_handle* h, int np, cudaStream_t st) {
    if (np <= h->cap_np) return 0;
    // grow geometrically past the first allocation so appends do not reallocate every time
    int cap = h->cap_np ? round_up(np + np / 4, PAD) : np;
    BO_CUDA(h, cudaStreamSynchronize(st));
    // keep the old state to preserve a fitted model across a growth (append path)
    bo_handle old = *h;
    h->Xs = h->Xraw = h->yv = h->alpha = h->Lm = h->Li = h->Tw = h->Lp = h->vec1 = h->vec2 = h->vec3 = nullptr;
    const size_t c = (size_t)cap;
    const size_t packed = (c / SW_BM) * (c / SW_BM + 1) / 2 * (SW_BM / SW_BK) * SW_TILE;
    struct { double** p; size_t elems; } reqs[] = {
        {&h->Xs, c * BO_MAX_DIM}, {&h->Xraw, c * BO_MAX_DIM}, {&h->yv, c}, {&h->alpha, c},
        {&h->Lm, c * c}, {&h->Li, c * c}, {&h->Tw, c * c / 4 + 16 * c + 64}, {&h->Lp, packed},
        {&h->vec1, c}, {&h->vec2, c}, {&h->vec3, c}};
    for (auto& r : reqs) {
        cudaError_t e = cudaMalloc(r.p, r.elems * sizeof(double));
        if (e != cudaSuccess) {
            for (auto& r2 : reqs) { if (*r2.p) cudaFree(*r2.p); *r2.p = nullptr; }
            h->Xs = old.Xs; h->Xraw = old.Xraw; h->yv = old.yv; h->alpha = old.alpha; h->Lm = old.Lm;
            h->Li = old.Li; h->Tw = old.Tw; h->Lp = old.Lp; h->vec1 = old.vec1; h->vec2 = old.vec2; h->vec3 = old.vec3;
            h->err = std::string("cudaMalloc failed growing capacity: ") + cudaGetErrorString(e);
            cudaGetLastError();
            return BO_E_NOMEM;
        }
    }
    h->cap_np = cap;
    h->plan_np = -1;
    if (old.cap_np > 0 && old.fitted) {
        // carry the fitted state over (row pitch changes from old.cap_np to cap)
        const size_t rows = old.np;
        BO_CUDA(h, cudaMemcpyAsync(h->Xs, old.Xs, rows * BO_MAX_DIM * 8, cudaMemcpyDeviceToDevice, st));
        BO_CUDA(h, cudaMemcpyAsync(h->Xraw, old.Xraw, rows * BO_MAX_DIM * 8, cudaMemcpyDeviceToDevice, st));
        BO_CUDA(h, cudaMemcpyAsync(h->yv, old.yv, rows * 8, cudaMemcpyDeviceToDevice, st));
        BO_CUDA(h, cudaMemcpyAsync(h->alpha, old.alpha, rows * 8, cudaMemcpyDeviceToDevice, st));
        BO_CUDA(h, cudaMemcpy2DAsync(h->Lm, c * 8, old.Lm, (size_t)old.cap_np * 8, rows * 8, rows, cudaMemcpyDeviceToDevice, st));
        BO_CUDA(h, cudaMemcpy2DAsync(h->Li, c * 8, old.Li, (size_t)old.cap_np * 8, rows * 8, rows, cudaMemcpyDeviceToDevice, st));
        BO_CUDA(h, cudaStreamSynchronize(st));
        // Lp (packed) does not depend on the pitch: tiles are indexed by (ib, kc) only
        const size_t oldpacked = ((size_t)old.np / SW_BM) * (old.np / SW_BM + 1) / 2 * (SW_BM / SW_BK) * SW_TILE;
        BO_CUDA(h, cudaMemcpy(h->Lp, old.Lp, oldpacked * 8, cudaMemcpyDeviceToDevice));
    }
    double* olds[] = {old.Xs, old.Xraw, old.yv, old.alpha, old.Lm, old.Li, old.Tw, old.Lp, old.vec1, old.vec2, old.vec3};
    for (double* p : olds) if (p) cudaFree(p);
    return 0;
}

template <int DP>
static int launch_gram(bo_handle* h, cudaStream_t st) {
    dim3 grid(h->np / 32, h->np / GRAM_ROWS), block(32, 8);
    gram_kernel<DP><<<grid, block, 0, st>>>(h->Xs, h->n, h->np, h->cap_np, h->hyp, h->Lm);
    BO_LAUNCH_CHECK(h);
    return 0;
}
#define BO_DISPATCH_DP(dp, fn, ...)                                   \
    ((dp) == 2 ? fn<2>(__VA_ARGS__) : (dp) == 4 ? fn<4>(__VA_ARGS__) :  \
     (dp) == 6 ? fn<6>(__VA_ARGS__) : (dp) == 8 ? fn<8>(__VA_ARGS__) :  \
     (dp) == 12 ? fn<12>(__VA_ARGS__) : fn<16>(__VA_ARGS__))

__global__ void sub_vec_kernel(const double* __restrict__ a, const double* __restrict__ b, int n, double* __restrict__ out) {
    a += (size_t)blockIdx.z * n; b += (size_t)blockIdx.z * n; out += (size_t)blockIdx.z * n;          // slot
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}

// alpha = (L L^T)^-1 r via the explicit inverse, plus one step of iterative refinement against the factor:
//   alpha0 = Li^T (Li r) ; r2 = r - L (L^T alpha0) ; alpha = alpha0 + Li^T (Li r2)
// (four HBM-bound triangular mat-vecs; makes alpha as accurate as a backward-stable cho_solve)
int solve_alpha_rhs(bo_handle* h, const double* r, double* alpha_out, cudaStream_t st) {
    const int np = h->np;
    int rc;
    if ((rc = trmv_lower_m(h, h->Li, r, h->vec2, st))) return rc;                                    // vec2 = Li r
    if ((rc = trmv_lower_t_m(h, h->Li, h->vec2, alpha_out, 0, st))) return rc;                       // alpha0
    if ((rc = trmv_lower_t_m(h, h->Lm, alpha_out, h->vec2, 0, st))) return rc;                       // vec2 = L^T alpha0
    if ((rc = trmv_lower_m(h, h->Lm, h->vec2, h->vec3, st))) return rc;                              // vec3 = L vec2
    sub_vec_kernel<<<(np + 255) / 256, 256, 0, st>>>(r, h->vec3, np, h->vec2);                       // vec2 = r2
    BO_LAUNCH_CHECK(h);
    if ((rc = trmv_lower_m(h, h->Li, h->vec2, h->vec3, st))) return rc;
    if ((rc = trmv_lower_t_m(h, h->Li, h->vec3, alpha_out, 1, st))) return rc;
    return 0;
}
static int solve_alpha(bo_handle* h, cudaStream_t st) {
    const int np = h->np;
    resid_init_kernel<<<(np + 255) / 256, 256, 0, st>>>(h->yv, h->n, np, h->hyp.mean, h->vec1);      // vec1 = r
    BO_LAUNCH_CHECK(h);
    return solve_alpha_rhs(h, h->vec1, h->alpha, st);
}

static int factor_and_pack(bo_handle* h, bool with_alpha, cudaStream_t st) {
    const int np = h->np, ld = h->cap_np, nb = np / NB;
    int rc;
    if (h->plan_np != np && (rc = build_plan(h, st))) return rc;
    BO_CUDA(h, cudaMemsetAsync(h->info_dev, 0, sizeof(int), st));
    if ((rc = BO_DISPATCH_DP(h->dp, launch_gram, h, st))) return rc;
    BO_CUDA(h, cudaMemset2DAsync(h->Li, (size_t)ld * 8, 0, (size_t)np * 8, np, st));
    for (int kb = 0; kb < nb; ++kb) {
        double* D = h->Lm + (size_t)kb * NB * ld + kb * NB;
        chol_panel_kernel<<<nb - kb, 256, 0, st>>>(D, ld, h->Li + (size_t)kb * NB * ld + kb * NB, h->info_dev, kb * NB, 0, nb - kb, chol_panel_fused(), chol_panel_kpre(kb));
        BO_LAUNCH_CHECK(h);
        if (kb + 1 < nb && (rc = gemm_launch(h, h->plan_launches[kb], st))) return rc;
    }
    BO_CUDA(h, cudaMemcpyAsync(h->info_host, h->info_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    leaf_inverse_kernel<<<nb, 256, 0, st>>>(h->Lm, ld, h->Li, 0, nb);
    BO_LAUNCH_CHECK(h);
    for (size_t li = (size_t)(nb - 1); li < h->plan_launches.size(); ++li)
        if ((rc = gemm_launch(h, h->plan_launches[li], st))) return rc;
    {
        dim3 grid(np / SW_BK, np / SW_BM);
        pack_linv_kernel<<<grid, 256, 0, st>>>(h->Li, ld, np, h->Lp, 0);
        BO_LAUNCH_CHECK(h);
    }
    if (with_alpha && (rc = solve_alpha(h, st))) return rc;
    BO_CUDA(h, cudaStreamSynchronize(st));
    const int info = *h->info_host;
    if (info != 0) {
        h->fitted = false;
        char buf[128];
        snprintf(buf, sizeof buf, "matrix not positive definite at pivot %d", info);
        h->err = buf;
        return info > h->n ? h->n : info;
    }
    h->fitted = true;
    h->factor_epoch++;
    return 0;
}
int refit_factor(bo_handle* h, cudaStream_t st) { return factor_and_pack(h, true, st); }

static int trmv_lower_m(bo_handle* h, const double* M, const double* v, double* z, cudaStream_t st) {
    trmv_lower_kernel<<<h->np / 8, 256, 0, st>>>(M, h->cap_np, h->np, v, z, 0);
    BO_LAUNCH_CHECK(h);
    return 0;
}
static int trmv_lower_t_m(bo_handle* h, const double* M, const double* z, double* out, int accumulate, cudaStream_t st) {
    trmv_lower_t_kernel<<<dim3(h->np / 32, TRMVT_SPLITS), 256, 0, st>>>(M, h->cap_np, h->np, z, h->Tw, 0);
    BO_LAUNCH_CHECK(h);
    trmv_reduce_kernel<<<(h->np + 255) / 256, 256, 0, st>>>(h->Tw, h->np, out, accumulate);
    BO_LAUNCH_CHECK(h);
    return 0;
}
int launch_trmv_lower(bo_handle* h, const double* v, double* z, cudaStream_t st) { return trmv_lower_m(h, h->Li, v, z, st); }
int launch_trmv_lower_t(bo_handle* h, const double* z, double* out, int accumulate, cudaStream_t st) {
    return trmv_lower_t_m(h, h->Li, z, out, accumulate, st);
}
// the row-split partial sums of L^-T z only (h->Tw: [splits][np]); the caller adds the splits in ascending order itself
// (bo_append folds that reduction into its finalize kernel: one launch less on a chain of launch latencies)
int launch_trmv_lower_t_partial(bo_handle* h, const double* z, const double** part, int* splits, cudaStream_t st) {
    trmv_lower_t_kernel<<<dim3(h->np / 32, TRMVT_SPLITS), 256, 0, st>>>(h->Li, h->cap_np, h->np, z, h->Tw, 0);
    BO_LAUNCH_CHECK(h);
    *part = h->Tw; *splits = TRMVT_SPLITS;
    return 0;
}

int pack_row_block(bo_handle* h, int ib, cudaStream_t st) {
    dim3 grid(h->np / SW_BK, 1);
    pack_linv_kernel<<<grid, 256, 0, st>>>(h->Li, h->cap_np, h->np, h->Lp, ib);
    BO_LAUNCH_CHECK(h);
    return 0;
}

int export_state(bo_handle* h, double* alpha_dev, double* chol_dev, double* linv_dev, cudaStream_t st) {
    BO_CUDA(h, cudaSetDevice(h->device));
    const int n = h->n;
    if (alpha_dev) BO_CUDA(h, cudaMemcpyAsync(alpha_dev, h->alpha, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    const unsigned blocks = (unsigned)(((size_t)n * n + 255) / 256);
    if (chol_dev) { export_lower_kernel<<<blocks, 256, 0, st>>>(h->Lm, h->cap_np, n, chol_dev); BO_LAUNCH_CHECK(h); }
    if (linv_dev) { export_lower_kernel<<<blocks, 256, 0, st>>>(h->Li, h->cap_np, n, linv_dev); BO_LAUNCH_CHECK(h); }
    return 0;
}

static int fit_impl_prepare(bo_handle* h, const double* X_dev, const double* y_dev, int n, int d, int kind,
             const double* ls_host, double outputscale, double noise, double mean, double jitter,
             double linear_variance, cudaStream_t st) {
    if (n < 1 || d < 1 || !X_dev || !y_dev || !ls_host) return fail(h, BO_E_INVALID, "bo_fit: bad argument");
    if (d > BO_MAX_DIM) return fail(h, BO_E_CAPACITY, "bo_fit: d exceeds BO_MAX_DIM");
    if (kind != BO_KERNEL_MATERN52 && kind != BO_KERNEL_RBF && kind != BO_KERNEL_LINEAR_MATERN52)
        return fail(h, BO_E_INVALID, "bo_fit: unknown kernel kind");
    if (kind == BO_KERNEL_LINEAR_MATERN52 && !(linear_variance >= 0.0)) return fail(h, BO_E_INVALID, "bo_fit: linear variance must be >= 0");
    if (!(outputscale > 0.0) || !(noise >= 0.0) || !(jitter >= 0.0)) return fail(h, BO_E_INVALID, "bo_fit: bad hyper-parameter");
    for (int k = 0; k < d; ++k) if (!(ls_host[k] > 0.0)) return fail(h, BO_E_INVALID, "bo_fit: lengthscale must be positive");
    BO_CUDA(h, cudaSetDevice(h->device));
    h->fitted = false;
    h->svgp = false;
    const int np = round_up(n, PAD);
    int rc = ensure_capacity(h, np, st);
    if (rc) return rc;
    h->n = n; h->np = np; h->d = d; h->dp = pad_dim(d);
    Hyper& hy = h->hyp;
    hy.kind = kind; hy.d = d; hy.dp = h->dp; hy.outputscale = outputscale; hy.noise = noise; hy.mean = mean; hy.jitter = jitter;
    hy.lin_v = kind == BO_KERNEL_LINEAR_MATERN52 ? linear_variance : 0.0;
    // gpytorch LinearKernel(ard_num_dims = d) carries one variance per input dimension (raw_variance (..., 1, d),
    // optimization/Bayesian7.py:162-166): v <x, x'> becomes sum_k v_k x_k x'_k -- on the scaled inputs the weights v_k l_k^2
    const bool ard_v = kind == BO_KERNEL_LINEAR_MATERN52 && (int)h->lin_v_ard.size() == d;
    if (kind == BO_KERNEL_LINEAR_MATERN52 && !h->lin_v_ard.empty() && !ard_v) {
        h->lin_v_ard.clear();
        return fail(h, BO_E_INVALID, "bo_set_linear_variance_ard: the number of variances does not match the input dimension of this fit");
    }
    for (int k = 0; k < BO_MAX_DIM; ++k) {
        hy.inv_ls[k] = k < d ? 1.0 / ls_host[k] : 0.0;
        hy.lin_w[k] = k < d ? (ard_v ? h->lin_v_ard[k] : hy.lin_v) * ls_host[k] * ls_host[k] : 0.0;
    }
    if (ard_v) {                                  // lin_v keeps the mean (introspection only); consumed by this fit
        double sv = 0.0; for (double v : h->lin_v_ard) sv += v;
        hy.lin_v = sv / d;
        h->lin_v_ard.clear();
    }
    stage_inputs_kernel<<<(np + 127) / 128, 128, 0, st>>>(X_dev, y_dev, n, d, np, hy, h->Xraw, h->Xs, h->yv);
    BO_LAUNCH_CHECK(h);
    return 0;
}
int fit_impl(bo_handle* h, const double* X_dev, const double* y_dev, int n, int d, int kind,
             const double* ls_host, double outputscale, double noise, double mean, double jitter,
             double linear_variance, cudaStream_t st) {
    const int rc = fit_impl_prepare(h, X_dev, y_dev, n, d, kind, ls_host, outputscale, noise, mean, jitter, linear_variance, st);
    return rc ? rc : refit_factor(h, st);
}

// ------------------------------------------------------------------------------------------
// N2: SVGP predictive state (whitened VariationalStrategy + CholeskyVariationalDistribution, Bayesian7.py:129-195)
// ------------------------------------------------------------------------------------------
// second triangular factor of the predictive variance: A2 = J Ls^T J (lower triangular in the reversed index order),
// packed exactly like L^-1 (pack_linv_kernel) so the sweep streams it with the same tile loop
__global__ void __launch_bounds__(256) pack_rev_chol_kernel(const double* __restrict__ Ls, int M, int np, double* __restrict__ Lp2) {
    constexpr int KCH = SW_BM / SW_BK;
    const int ib = blockIdx.y;
    const int kc = blockIdx.x;
    if (kc >= (ib + 1) * KCH) return;
    double* dst = Lp2 + ((size_t)ib * (ib + 1) / 2 * KCH + kc) * SW_TILE;
    const int row0 = ib * SW_BM, k0 = kc * SW_BK;
    for (int e = threadIdx.x; e < SW_TILE; e += 256) {
        const int k = e / SW_BM, r = e % SW_BM;          // consecutive threads walk a row of Ls backwards
        const int ip = row0 + r, jp = k0 + k;            // A2[ip][jp] = Ls[np-1-jp][np-1-ip]
        const int sr = np - 1 - jp, sc = np - 1 - ip;
        const double v = (jp <= ip && sr < M) ? Ls[(size_t)sr * M + sc] : 0.0;
        const int r8 = r >> 3, k8 = k >> 3;
        const int lane = (r & 7) * 4 + (k & 3), khalf = (k & 7) >> 2;
        dst[(r8 * (SW_BK / 8) + k8) * 64 + lane * 2 + khalf] = v;
    }
}

// Ls (M x M, lower triangle used) -> zero-padded np x np copy, so that the GEMM below reads whole tiles
__global__ void pad_lower_kernel(const double* __restrict__ Ls, int M, int np, double* __restrict__ out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)np * np) return;
    const int i = (int)(e / np), j = (int)(e % np);
    out[e] = (i < M && j <= i) ? Ls[(size_t)i * M + j] : 0.0;
}

// B = Ls^T L^-1 (dense np x np) for the sliced sweep of an SVGP state: ||Ls^T u||^2 = ||B k*||^2 makes the second term of the
// predictive variance a contraction of the SAME sliced panel k* as ||L^-1 k*||^2 -- one pass over the stacked factor [L^-1; B]
// instead of a second pass over u (which would have to be re-sliced per candidate).  One grouped-GEMM launch at load time.
static int svgp_build_b(bo_handle* h, const double* Ls_dev, int M, cudaStream_t st) {
    const int np = h->np;
    const size_t elems = (size_t)np * np;
    if (elems > h->svB_elems) {
        if (h->svB) cudaFree(h->svB);
        h->svB = nullptr; h->svB_elems = 0;
        BO_CUDA(h, cudaMalloc(&h->svB, elems * sizeof(double)));
        h->svB_elems = elems;
    }
    double* pad = nullptr; GemmProblem* pd = nullptr;
    BO_CUDA(h, cudaMalloc(&pad, elems * sizeof(double)));
    if (cudaMalloc(&pd, sizeof(GemmProblem)) != cudaSuccess) { cudaFree(pad); cudaGetLastError(); return fail(h, BO_E_NOMEM, "bo_svgp_load: allocation failed"); }
    pad_lower_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(Ls_dev, M, np, pad);
    h->launches++;
    GemmBatch g(64);
    // C[m][n] = sum_k Ls[k][m] L^-1[k][n]: A given as [K][M]; both operands vanish for k < max(m, n)
    g.add(pad, np, h->Li, h->cap_np, h->svB, np, np, np, np, 1.0, 0.0, /*transB=*/0, GEMM_TRANS_A | GEMM_K_FROM_MAX);
    cudaError_t e = cudaMemcpyAsync(pd, g.probs.data(), sizeof(GemmProblem), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        dgemm_grouped_kernel<64, 64, 16, 2, 4><<<g.tiles, 256, GemmSmem<64, 64, 16, 2>::BYTES, st>>>(pd, 1);
        h->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(pad); cudaFree(pd);
    if (e != cudaSuccess) { h->err = std::string("bo_svgp_load: building Ls^T L^-1 failed: ") + cudaGetErrorString(e); cudaGetLastError(); return BO_E_CUDA; }
    return 0;
}

int svgp_load_impl(bo_handle* h, const double* Z_dev, int M, int d, int kind, const double* ls_host, double outputscale,
                   double linear_variance, double mean, double noise, double jitter, const double* var_mean_dev,
                   const double* var_chol_dev, cudaStream_t st) {
    if (M < 1 || d < 1 || !Z_dev || !ls_host || !var_mean_dev || !var_chol_dev) return fail(h, BO_E_INVALID, "bo_svgp_load: bad argument");
    if (!(noise >= 0.0)) return fail(h, BO_E_INVALID, "bo_svgp_load: noise must be >= 0");
    // inducing-point prior: K_uu + jitter I (no likelihood noise inside the factorisation); m rides in as the "targets"
    int rc = fit_impl_prepare(h, Z_dev, var_mean_dev, M, d, kind, ls_host, outputscale, 0.0, mean, jitter, linear_variance, st);
    if (rc) return rc;
    if ((rc = factor_and_pack(h, false, st))) return rc;
    h->fitted = false;                                   // until the second factor is in place
    const int np = h->np;
    if ((rc = trmv_lower_t_m(h, h->Li, h->yv, h->alpha, 0, st))) return rc;          // alpha = L^-T m  (mean = c + k*^T alpha)
    const size_t packed = ((size_t)np / SW_BM) * (np / SW_BM + 1) / 2 * (SW_BM / SW_BK) * SW_TILE;
    if (packed > h->Lp2_elems) {
        if (h->Lp2) cudaFree(h->Lp2);
        h->Lp2 = nullptr; h->Lp2_elems = 0;
        BO_CUDA(h, cudaMalloc(&h->Lp2, packed * sizeof(double)));
        h->Lp2_elems = packed;
    }
    pack_rev_chol_kernel<<<dim3(np / SW_BK, np / SW_BM), 256, 0, st>>>(var_chol_dev, M, np, h->Lp2);
    BO_LAUNCH_CHECK(h);
    if ((rc = svgp_build_b(h, var_chol_dev, M, st))) return rc;
    BO_CUDA(h, cudaStreamSynchronize(st));               // var_chol_dev is borrowed for the call only
    h->svgp = true;
    h->sv_add = jitter + noise;
    h->fitted = true;
    return 0;
}

// GEMM throughput probe (development / roofline evidence for the fit's trailing updates): C = A * B^T, square tiles
int gemm_probe_impl(bo_handle* h, int m, int n, int k, int cfg, int reps, double* tflops) {
    BO_CUDA(h, cudaSetDevice(h->device));
    const int form = cfg / 10;          // 0: A [M][K], B [N][K] (SYRK / TRSM shape) | 1: B [K][N] (inverse levels) | 2: A [K][M], B [K][N] (L^-T L^-1)
    cfg %= 10;
    const int bm = (cfg == 1 || cfg == 2 || cfg == 6) ? 128 : 64, bn = (cfg == 1 || cfg == 6) ? 128 : 64;
    {   // probe-only variants opt in here
        cudaFuncSetAttribute(dgemm_grouped_kernel<64, 64, 16, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmSmem<64, 64, 16, 2>::BYTES);
        cudaFuncSetAttribute(dgemm_grouped_kernel<64, 64, 32, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmSmem<64, 64, 32, 2>::BYTES);
        cudaFuncSetAttribute(dgemm_grouped_kernel<64, 64, 32, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmSmem<64, 64, 32, 3>::BYTES);
        cudaFuncSetAttribute(dgemm_grouped_kernel<128, 128, 32, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmSmem<128, 128, 32, 2>::BYTES);
    }
    if (m % bm || n % bn || k % 32 || m < bm || n < bn || k < 16) return fail(h, BO_E_INVALID, "bo_gemm_probe: sizes must be tile multiples");
    double *A, *B, *C; GemmProblem* pd;
    BO_CUDA(h, cudaMalloc(&A, (size_t)m * k * 8)); BO_CUDA(h, cudaMalloc(&B, (size_t)n * k * 8)); BO_CUDA(h, cudaMalloc(&C, (size_t)m * n * 8));
    BO_CUDA(h, cudaMalloc(&pd, sizeof(GemmProblem)));
    BO_CUDA(h, cudaMemset(A, 0, (size_t)m * k * 8)); BO_CUDA(h, cudaMemset(B, 0, (size_t)n * k * 8)); BO_CUDA(h, cudaMemset(C, 0, (size_t)m * n * 8));
    GemmProblem p{}; p.A = A; p.B = B; p.C = C; p.M = m; p.N = n; p.K = k; p.lda = k; p.ldb = k; p.ldc = n; p.alpha = 1.0; p.beta = 1.0;
    p.transB = 1; p.mode = 0; p.tiles_n = n / bn; p.tile_begin = 0; p.tile_end = (m / bm) * (n / bn);
    if (form >= 1) { p.transB = 0; p.ldb = n; }
    if (form == 2) { p.mode = GEMM_TRANS_A; p.lda = m; }
    BO_CUDA(h, cudaMemcpy(pd, &p, sizeof p, cudaMemcpyHostToDevice));
    auto launch = [&]() {
        if (cfg == 1) dgemm_grouped_kernel<128, 128><<<p.tile_end, 256, GemmSmem<128, 128>::BYTES>>>(pd, 1);
        else if (cfg == 2) dgemm_grouped_kernel<128, 64><<<p.tile_end, 256, GemmSmem<128, 64>::BYTES>>>(pd, 1);
        else if (cfg == 3) dgemm_grouped_kernel<64, 64, 16, 2, 4><<<p.tile_end, 256, GemmSmem<64, 64, 16, 2>::BYTES>>>(pd, 1);
        else if (cfg == 4) dgemm_grouped_kernel<64, 64, 32, 2, 3><<<p.tile_end, 256, GemmSmem<64, 64, 32, 2>::BYTES>>>(pd, 1);
        else if (cfg == 5) dgemm_grouped_kernel<64, 64, 32, 3, 2><<<p.tile_end, 256, GemmSmem<64, 64, 32, 3>::BYTES>>>(pd, 1);
        else if (cfg == 6) dgemm_grouped_kernel<128, 128, 32, 2, 1><<<p.tile_end, 256, GemmSmem<128, 128, 32, 2>::BYTES>>>(pd, 1);
        else dgemm_grouped_kernel<64, 64><<<p.tile_end, 256, GemmSmem<64, 64>::BYTES>>>(pd, 1);
        h->launches++;
    };
    launch(); BO_CUDA(h, cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    BO_CUDA(h, cudaGetLastError());
    *tflops = 2.0 * m * n * k * reps / (ms * 1e-3) * 1e-12;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(pd);
    return 0;
}

#include "lml.cuh"

}  // namespace bo
