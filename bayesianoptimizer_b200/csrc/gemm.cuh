// Grouped FP64 GEMM on the DMMA tensor path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4) for the fit:
// Cholesky panel/trailing updates and the recursive triangular inverse.  One launch executes a
// list of independent problems  C = alpha * A * op(B) + beta * C  (row-major), flattened over tiles.
// Operands are staged with cp.async (LDGSTS) into a 3-stage shared-memory ring.
#pragma once
#include "common.cuh"

namespace bo {

enum : int {
    GEMM_LOWER_C    = 1,   // skip output tiles strictly above the block diagonal (SYRK, lower)
    GEMM_A_LOWER    = 2,   // A (M x K, square) is lower triangular: k < (tile_m + 1) * BM
    GEMM_B_LOWER_NN = 4,   // B (K x N, square) is lower triangular: k >= tile_n * BN
    GEMM_B_LOWER_NT = 8,   // B given as [N][K] lower triangular (k <= n): k < (tile_n + 1) * BN
    GEMM_TRANS_A    = 16,  // A given as [K][M] (column access): C = alpha * A^T * op(B) + beta * C
    GEMM_K_FROM_MAX = 32,  // both operands vanish for k < max(tile_m * BM, tile_n * BN) (L^-T L^-1 products)
};

constexpr int GEMM_BK = 16;
constexpr int GEMM_STAGES = 3;
constexpr int GEMM_LDK = GEMM_BK + 4;     // [row][k] layout stride (conflict-free DMMA fragment reads)

template <int BM, int BN>
struct GemmSmem {
    static constexpr int A_N = BM * GEMM_LDK;
    static constexpr int A_T = GEMM_BK * (BM + 4);
    static constexpr int A_ELEMS = A_N > A_T ? A_N : A_T;
    static constexpr int B_NT = BN * GEMM_LDK;
    static constexpr int B_NN = GEMM_BK * (BN + 4);
    static constexpr int B_ELEMS = B_NT > B_NN ? B_NT : B_NN;
    static constexpr int STAGE = A_ELEMS + B_ELEMS;
    static constexpr size_t BYTES = (size_t)GEMM_STAGES * STAGE * sizeof(double);
};

template <int BM, int BN>
__global__ void __launch_bounds__(256) dgemm_grouped_kernel(const GemmProblem* __restrict__ probs, int nprob, int tile0 = 0) {
    extern __shared__ __align__(16) double gsm[];
    using SM = GemmSmem<BM, BN>;
    constexpr int WM = BM / 2, WN = BN / 4;      // 8 warps as 2 (M) x 4 (N)
    constexpr int MT = WM / 8, NT = WN / 8;

    // locate the problem of this tile
    const int tile = blockIdx.x + tile0;     // tile0: first flattened tile when `probs` is a sub-range of a launch list
    int pi = 0;
    while (pi + 1 < nprob && tile >= probs[pi].tile_end) ++pi;
    const GemmProblem P = probs[pi];
    const int lt = tile - P.tile_begin;
    const int tm = lt / P.tiles_n, tn = lt % P.tiles_n;
    if ((P.mode & GEMM_LOWER_C) && tn * BN > tm * BM + (BM - 1)) return;

    int k_begin = 0, k_end = P.K;
    if (P.mode & GEMM_A_LOWER)    k_end = min(k_end, (tm + 1) * BM);
    if (P.mode & GEMM_B_LOWER_NN) k_begin = max(k_begin, tn * BN);
    if (P.mode & GEMM_B_LOWER_NT) k_end = min(k_end, (tn + 1) * BN);
    if (P.mode & GEMM_K_FROM_MAX) k_begin = max(k_begin, max(tm * BM, tn * BN));
    const bool transA = (P.mode & GEMM_TRANS_A) != 0;
    const int nk = (k_end - k_begin) / GEMM_BK;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, q = lane & 3;

    const double* Ag = transA ? P.A + tm * BM : P.A + (size_t)(tm * BM) * P.lda;
    const double* Bg = P.transB ? P.B + (size_t)(tn * BN) * P.ldb : P.B + tn * BN;

    auto load_stage = [&](int s, int kt) {
        double* As = gsm + s * SM::STAGE;
        double* Bs = As + SM::A_ELEMS;
        const int k0 = k_begin + kt * GEMM_BK;
        if (transA) {
#pragma unroll
            for (int c = tid; c < GEMM_BK * (BM / 2); c += 256) {
                int kr = c / (BM / 2), mq = c % (BM / 2);
                cp_async16(As + kr * (BM + 4) + mq * 2, Ag + (size_t)(k0 + kr) * P.lda + mq * 2);
            }
        } else {
#pragma unroll
            for (int c = tid; c < BM * (GEMM_BK / 2); c += 256) {
                int row = c >> 3, kq = c & 7;
                cp_async16(As + row * GEMM_LDK + kq * 2, Ag + (size_t)row * P.lda + k0 + kq * 2);
            }
        }
        if (P.transB) {
#pragma unroll
            for (int c = tid; c < BN * (GEMM_BK / 2); c += 256) {
                int row = c >> 3, kq = c & 7;
                cp_async16(Bs + row * GEMM_LDK + kq * 2, Bg + (size_t)row * P.ldb + k0 + kq * 2);
            }
        } else {
#pragma unroll
            for (int c = tid; c < GEMM_BK * (BN / 2); c += 256) {
                int kr = c / (BN / 2), nq = c % (BN / 2);
                cp_async16(Bs + kr * (BN + 4) + nq * 2, Bg + (size_t)(k0 + kr) * P.ldb + nq * 2);
            }
        }
    };

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<GEMM_STAGES - 2>();
        __syncthreads();
        {   // prefetch tile kt + STAGES - 1 into the slot consumed at iteration kt - 1
            int nx = kt + GEMM_STAGES - 1;
            if (nx < nk) load_stage(nx % GEMM_STAGES, nx);
            cp_async_commit();
        }
        const double* As = gsm + (kt % GEMM_STAGES) * SM::STAGE;
        const double* Bs = As + SM::A_ELEMS;
#pragma unroll
        for (int kk = 0; kk < GEMM_BK / 4; ++kk) {
            double a[MT], b[NT];
            if (transA) {
#pragma unroll
                for (int i = 0; i < MT; ++i) a[i] = As[(kk * 4 + q) * (BM + 4) + wm * WM + i * 8 + g];
            } else {
#pragma unroll
                for (int i = 0; i < MT; ++i) a[i] = As[(wm * WM + i * 8 + g) * GEMM_LDK + kk * 4 + q];
            }
            if (P.transB) {
#pragma unroll
                for (int j = 0; j < NT; ++j) b[j] = Bs[(wn * WN + j * 8 + g) * GEMM_LDK + kk * 4 + q];
            } else {
#pragma unroll
                for (int j = 0; j < NT; ++j) b[j] = Bs[(kk * 4 + q) * (BN + 4) + wn * WN + j * 8 + g];
            }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // epilogue: each thread owns (row g, cols 2q, 2q+1) of every 8x8 tile
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int row = tm * BM + wm * WM + i * 8 + g;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int col = tn * BN + wn * WN + j * 8 + 2 * q;
            double2* cp = reinterpret_cast<double2*>(P.C + (size_t)row * P.ldc + col);
            double2 o;
            if (P.beta != 0.0) {
                double2 old = *cp;
                o.x = fma(P.alpha, acc[i][j][0], P.beta * old.x);
                o.y = fma(P.alpha, acc[i][j][1], P.beta * old.y);
            } else {
                o.x = P.alpha * acc[i][j][0];
                o.y = P.alpha * acc[i][j][1];
            }
            *cp = o;
        }
    }
}

// Host-side plan builder: appends one launch (a list of problems sharing a tile shape).
struct GemmBatch {
    std::vector<GemmProblem> probs;
    int tiles = 0;
    int bm = 64;
    explicit GemmBatch(int tile) : bm(tile) {}
    void add(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K,
             double alpha, double beta, int transB, int mode) {
        if (M <= 0 || N <= 0 || K <= 0) return;
        GemmProblem p{};
        p.A = A; p.B = B; p.C = C; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
        p.alpha = alpha; p.beta = beta; p.transB = transB; p.mode = mode;
        p.tiles_n = N / bm;
        p.tile_begin = tiles;
        tiles += (M / bm) * (N / bm);
        p.tile_end = tiles;
        probs.push_back(p);
    }
};

int gemm_init(bo_handle* h);
// launch `count` problems starting at plan_dev + first
int gemm_launch(bo_handle* h, const GemmLaunch& L, cudaStream_t st);

}  // namespace bo
