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

// Tile shape BM x BN, contraction chunk BK per stage, STAGES-deep cp.async ring, MINB resident CTAs per SM requested.
template <int BM, int BN, int BK = 16, int STAGES = 3>
struct GemmSmem {
    static constexpr int LDK = BK + 4;        // [row][k] layout stride (conflict-free DMMA fragment reads)
    static constexpr int A_N = BM * LDK;
    static constexpr int A_T = BK * (BM + 4);
    static constexpr int A_ELEMS = A_N > A_T ? A_N : A_T;
    static constexpr int B_NT = BN * LDK;
    static constexpr int B_NN = BK * (BN + 4);
    static constexpr int B_ELEMS = B_NT > B_NN ? B_NT : B_NN;
    static constexpr int STAGE = A_ELEMS + B_ELEMS;
    static constexpr size_t BYTES = (size_t)STAGES * STAGE * sizeof(double);
};

template <int BM, int BN, int BK = 16, int STAGES = 3, int MINB = (BM * BN <= 4096 ? 3 : 1)>
__global__ void __launch_bounds__(256, MINB) dgemm_grouped_kernel(const GemmProblem* __restrict__ probs, int nprob, int tile0 = 0) {
    extern __shared__ __align__(16) double gsm[];
    using SM = GemmSmem<BM, BN, BK, STAGES>;
    constexpr int LDK = SM::LDK;
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
    const int nk = (k_end - k_begin) / BK;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, q = lane & 3;

    // per-thread copy assignment, fixed over the k loop: [row][k] operands move BK/2 16-byte chunks per row,
    // [k][col] operands BM/2 (BN/2) chunks per k row; only the k offset advances
    constexpr int CK = BK / 2;                       // chunks per row, [row][k] form
    constexpr int RS = 256 / CK;                     // rows covered per pass
    const double* a_src; int a_dst; size_t a_step;   // source, smem offset, source step per pass
    const double* b_src; int b_dst; size_t b_step;
    if (transA) {
        const int mq = tid % (BM / 2), kr = tid / (BM / 2);
        a_src = P.A + tm * BM + (size_t)(k_begin + kr) * P.lda + mq * 2;
        a_dst = kr * (BM + 4) + mq * 2;
        a_step = (size_t)(256 / (BM / 2)) * P.lda;
    } else {
        const int row = tid / CK, kq = tid % CK;
        a_src = P.A + (size_t)(tm * BM + row) * P.lda + k_begin + kq * 2;
        a_dst = row * LDK + kq * 2;
        a_step = (size_t)RS * P.lda;
    }
    if (P.transB) {
        const int row = tid / CK, kq = tid % CK;
        b_src = P.B + (size_t)(tn * BN + row) * P.ldb + k_begin + kq * 2;
        b_dst = row * LDK + kq * 2;
        b_step = (size_t)RS * P.ldb;
    } else {
        const int nq = tid % (BN / 2), kr = tid / (BN / 2);
        b_src = P.B + tn * BN + (size_t)(k_begin + kr) * P.ldb + nq * 2;
        b_dst = kr * (BN + 4) + nq * 2;
        b_step = (size_t)(256 / (BN / 2)) * P.ldb;
    }
    const size_t a_kadv = transA ? (size_t)BK * P.lda : (size_t)BK;       // source advance per k tile
    const size_t b_kadv = P.transB ? (size_t)BK : (size_t)BK * P.ldb;

    auto load_stage = [&](int s, int kt) {
        double* As = gsm + s * SM::STAGE;
        double* Bs = As + SM::A_ELEMS;
        const double* ap = a_src + (size_t)kt * a_kadv;
        const double* bp = b_src + (size_t)kt * b_kadv;
        if (transA) {
#pragma unroll
            for (int it = 0; it < BK * (BM / 2) / 256; ++it) cp_async16(As + a_dst + it * (256 / (BM / 2)) * (BM + 4), ap + it * a_step);
        } else {
#pragma unroll
            for (int it = 0; it < BM * CK / 256; ++it) cp_async16(As + a_dst + it * RS * LDK, ap + it * a_step);
        }
        if (P.transB) {
#pragma unroll
            for (int it = 0; it < BN * CK / 256; ++it) cp_async16(Bs + b_dst + it * RS * LDK, bp + it * b_step);
        } else {
#pragma unroll
            for (int it = 0; it < BK * (BN / 2) / 256; ++it) cp_async16(Bs + b_dst + it * (256 / (BN / 2)) * (BN + 4), bp + it * b_step);
        }
    };

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    // fragment read offsets (fixed over the k loop)
    const int a_frag = transA ? q * (BM + 4) + wm * WM + g : (wm * WM + g) * LDK + q;
    const int b_frag = P.transB ? (wn * WN + g) * LDK + q : q * (BN + 4) + wn * WN + g;
    const int a_i = transA ? 8 : 8 * LDK, a_k = transA ? 4 * (BM + 4) : 4;      // steps per 8-row group / per k4 step
    const int b_j = P.transB ? 8 * LDK : 8, b_k = P.transB ? 4 : 4 * (BN + 4);
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {   // prefetch tile kt + STAGES - 1 into the slot consumed at iteration kt - 1
            int nx = kt + STAGES - 1;
            if (nx < nk) load_stage(nx % STAGES, nx);
            cp_async_commit();
        }
        const double* As = gsm + (kt % STAGES) * SM::STAGE + a_frag;
        const double* Bs = gsm + (kt % STAGES) * SM::STAGE + SM::A_ELEMS + b_frag;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[MT], b[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) a[i] = As[i * a_i + kk * a_k];
#pragma unroll
            for (int j = 0; j < NT; ++j) b[j] = Bs[j * b_j + kk * b_k];
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
