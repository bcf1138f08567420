// K4-i8: the sweep's variance contraction u = L^-1 k* as an error-free (Ozaki scheme I) product of signed 7-bit
// slices on the INT8 tensor path (tcgen05.mma kind::i8, INT32 accumulators in TMEM) -- included by sweep.cu.
//
// Why: the FP64 sweep (sweep_kernel) sits at 0.95 of the DMMA roof (37 TFLOP/s); INT8 tcgen05 runs at 4.4 POP/s
// (tools/i8_probe.cu).  Every row of L^-1 is scaled by a power of two to |x| < 1 and cut into S balanced digits (I8Dig:
// eight 7-bit ones, or one 7-bit + six 8-bit ones); the candidates' K(X, x*) columns likewise (fixed scale, the
// kernel is bounded by the output scale).  Digit products are exact integers; products with equal s + t share one
// INT32 accumulator (|sum| <= S * 64 * 64 * n < 2^31 for n < 65 536 at S = 8: the eligibility bound); the S accumulators are recombined in FP64
// (Horner in 2^-7) when a 128-row block is complete, squared and summed per candidate.  Pairs with s + t >= S are
// dropped: S (S + 1) / 2 products per FP64 multiply-add.  tools/ozaki_feasibility.py: S = 7 reproduces sigma^2 to
// 6e-10 relative at the C3 shape (bar: 1e-8), S = 8 is indistinguishable from the FP64 product.
//
// Shape: TMEM holds 512 columns, so S accumulators allow N = 64 candidates per CTA tile (M = 128 rows of L^-1).
// One persistent CTA of 384 threads per SM, three warpgroups (registers re-balanced with setmaxnreg):
//   builders (warps 8-11)  for the NEXT block of 64 candidates: Sobol point / explicit row, K(X, X*) in FP64, posterior
//            mean, digits (integer shifts and masks) -> one of this CTA's two int8 panel buffers in HBM/L2;
//   warp 4   streams stage tiles (S slices of a 128 x 64 piece of L^-1 and of a 64 x 64 piece of the current panel,
//            both pre-arranged in the tensor core's K-major 8 x 16 B core-matrix order) with TMA bulk copies into a
//            2-stage ring;
//   warp 5   one elected lane issues the S (S + 1) / 2 x 2 MMAs of a stage (fully unrolled, descriptors in uniform
//            registers) and commits to the ring's empty barrier and, per row block, to the accumulators-full barrier;
//   warps 0-3 (one TMEM lane = one row of L^-1 each) drain the accumulators with tcgen05.ld, recombine, keep
//            sum_i u_i^2 per candidate in registers, then run the epilogue: variance, acquisition, CTA-local top-k
//            (as sweep_kernel).
// Replaces the same reference code as sweep_kernel (optimization/Bayesian7.py:664-682, Bayesian.py:105-112).

constexpr int I8_BN      = 64;                 // candidates per CTA tile (TMEM: S * 64 columns)
#ifndef BO_I8_KC
#define BO_I8_KC 64
#endif
#ifndef BO_I8_STAGES
#define BO_I8_STAGES (BO_I8_KC == 32 ? 4 : 2)
#endif
#ifndef BO_P8_STAGES
#define BO_P8_STAGES (BO_I8_KC == 32 ? 5 : 2)      // CTA-pair kernel (sweep_i8_pair.cuh): 40 KB stages at K = 32
#endif
constexpr int I8_KC      = BO_I8_KC;           // contraction bytes per pipeline stage
constexpr int I8_STAGES  = BO_I8_STAGES;
constexpr int I8_THREADS = 384;                // warpgroup 0 (warps 0-3): drain + epilogue; 1: warp 4 TMA, warp 5 MMA issue, 6-7 idle; 2 (warps 8-11): panel builders
constexpr int I8_A_SLICE = SW_BM * I8_KC;      // 8 KB
constexpr int I8_B_SLICE = I8_BN * I8_KC;      // 4 KB
constexpr int I8_MIN_NP  = 256;                // below this the stage pipeline is all start-up (pinned modes fall back to FP64)
constexpr int I8_PAIR_MIN_NP = 2048;           // from here up the CTA-pair kernel (sweep_i8_pair.cuh) is the faster one
constexpr int I8_AUTO_MIN_NP = 512;            // AUTO: smallest padded n the sliced path was measured faster at

template <int S, int DP>
struct I8Smem {
    static constexpr int STAGE_BYTES = S * (I8_A_SLICE + I8_B_SLICE);
    static constexpr int OFF_BAR   = I8_STAGES * STAGE_BYTES;           // full[2], empty[2], tmem full/empty, panel full[2]/empty[2], tmem base
    static constexpr int OFF_COL   = OFF_BAR + 128;                     // colsum[4][64]
    static constexpr int OFF_MU    = OFF_COL + 4 * I8_BN * 8;           // mu[2][64]
    static constexpr int OFF_TKV   = OFF_MU + 2 * I8_BN * 8;
    static constexpr int OFF_TKI   = OFF_TKV + BO_MAX_TOPK * 8;
    static constexpr int OFF_ACQ   = OFF_TKI + BO_MAX_TOPK * 8;
    static constexpr int OFF_CMASK = OFF_ACQ + I8_BN * 8;
    static constexpr int OFF_SOB   = OFF_CMASK + 32;
    static constexpr int OFF_TAB   = (OFF_SOB + (BO_MAX_DIM * BO_SOBOL_BITS + BO_MAX_DIM) * 4 + 127) / 128 * 128;   // 2^(j/16), j < 16 (exp_neg_fast)
    static constexpr int OFF_X     = OFF_TAB + 128;
    // builders' X~ / alpha staging: two buffers of XCH rows, as many rows as the 227 KB budget leaves
    static constexpr int X_ROW     = (DP + 2 + 1) * 8;
    static constexpr int X_FREE    = 232448 - OFF_X;
    static constexpr int XCH       = X_FREE >= 2 * 512 * X_ROW ? 512 : X_FREE >= 2 * 256 * X_ROW ? 256 : X_FREE >= 2 * 128 * X_ROW ? 128 : 64;
    static_assert(X_FREE >= 2 * XCH * X_ROW, "no room for the X~ staging buffers");
    static constexpr int BYTES     = OFF_X + 2 * XCH * X_ROW;
};
struct SweepI8Args {
    const int8_t* Lp8; const double* rowscale; int8_t* panel8;
    const int8_t* Lp8_zero;  // one all-zero stage tile of L^-1 slices (CTA-pair kernel: K steps beyond a row block's own extent)
    double dig_scale;       // 2^(6 + 7 (S - 1)) / (power-of-two bound eb of |k*|): k* -> fixed point
    double eb_scale;        // 2^-12: folded into the row scale at the drain (the accumulators count units of 2^-12 eb rowscale)
    double ss_scale;        // eb^2: applied to the finished column sum (a power of two: exact, commutes with the summation).
                            // Linear + Matern kind: |k*| has no model-wide bound -- the CTA-pair kernel derives eb per
                            // candidate from s2 (sqrt(|x*|_w^2 guard_w[1]) + 1) and ignores dig_scale / ss_scale
    // accuracy guard: a candidate whose variance s2 - ||u||^2 is below guard_scale * sqrt(W ||u||^2), W = guard_w[0] =
    // max_i rowscale_i^2 (i + 1), is not scored here: its pool position goes to flag_idx (warp-aggregated append through
    // flag_count) and the FP64 contraction re-scores it after the kernel
    double guard_scale; const double* guard_w;
    long long* flag_idx; int* flag_count; long long flag_cap;
    // SVGP predictive state (CTA-pair kernel only): Lp8 holds the slices of the stack [L^-1; B], B = Ls^T L^-1 (dense row blocks
    // behind the triangular ones), rowscale has 2 np entries, guard_w[2] is W of B's rows, and the variance is
    // k** + sv_add - ||L^-1 k*||^2 + ||B k*||^2
    int sv; double sv_add;
};

// Slicing error of ||u||^2 (what the guard bounds).  The products of slices s + t >= S are dropped: per element product an
// error of about 2^GEXP rowscale_i eb with random sign (eb: the power-of-two bound of |k*|; GEXP = 2 W - 14 - W S from the digit
// geometry I8Dig: -56 for eight 7-bit slices, -54 for one 7-bit + six 8-bit ones), (S - 1) dropped pairs of rms 1/3 each,
// i + 1 products per row:
//     delta u_i ~ 2^GEXP eb rowscale_i sqrt((S - 1)(i + 1)) / 3,      delta ||u||^2 = 2 sum_i u_i delta u_i
//     rms(delta ||u||^2) <= 2 sqrt(S - 1) / 3  2^GEXP eb sqrt(W ||u||^2)          (< 1.8 x that expression for S <= 8).
// tools/ozaki_guard_study.py measures the true 7- and 8-slice errors on the reference's CSV rows against it.  A candidate is
// flagged when I8_GUARD_KAPPA times the expression exceeds I8_GUARD_RTOL of its variance: unflagged candidates carry a slicing
// error below 2e-9 relative (4.5 sigma), a fifth of the 1e-8 bar, on top of the FP64 recombination's own rounding.
constexpr double I8_GUARD_KAPPA = 8.0;
constexpr double I8_GUARD_RTOL  = 2e-9;

// Digit geometry.  An operand value x (|x| <= 1) becomes the integer v = rint(x 2^F) and is cut into S signed digits:
// a 7-bit top digit d[0] in [-64, 64] and S - 1 lower digits of W bits each, balanced ([-2^(W-1), 2^(W-1) - 1]):
//     x ~= d[0] 2^-6 + sum_{s >= 1} d[s] 2^(-6 - W s),      F = 6 + W (S - 1).
//   S = 8: W = 7 -> F = 55: eight 7-bit slices, 36 slice products per multiply-add (BO_SWEEP_I8X8);
//   S = 7: W = 8 -> F = 54: one 7-bit + six 8-bit slices, 28 products (BO_SWEEP_I8X7).  One bit less than the 8-slice form,
//          a 3.3x larger slicing error of ||u||^2 on the reference's CSV rows (CPU emulation: 2.6e-14 / 1.2e-14 / 4.4e-14
//          absolute at n = 512 / 512 / 3000 against 8.3e-15 / 3.1e-15 / 1.3e-14, FP64 BLAS 4.8e-15 / 3.7e-15 / 1.4e-14;
//          seven 7-BIT slices: 1.3e-12) -- for 22 % fewer MMAs.  INT32 accumulators: |sum| <= 7 * 128^2 * np < 2^31 up
//          to np = 16 384 (8 * 64^2 * np < 2^31 up to 65 535 for S = 8).
template <int S> struct I8Dig {
    static constexpr int W = (S == 7) ? 8 : 7;
    static constexpr int F = 6 + W * (S - 1);
    static constexpr int NP_MAX = (S == 7) ? 16384 : 65535;
    static constexpr double HORNER = (S == 7) ? 0.00390625 : 0.0078125;        // 2^-W: one accumulator group down
    // slicing-error scale of one element product: sqrt(S-1)/3 * 2^GEXP (dropped pairs s + t >= S: (S-1) products of two
    // digits of rms 2^(W-1)/sqrt(3) at weight 2^(-12 - W S)); GEXP = 2 W - 14 - W S  (= -7 S for W = 7)
    static constexpr int GEXP = 2 * W - 14 - W * S;
};
// the S digits of xs = x 2^F as plain bit fields: adding the bias sum_{k < S-1} 2^(W-1) 2^(W k) turns the balanced lower
// digits into the W-bit fields of one 64-bit integer (an exact decomposition: v = sum_k d_k 2^(W k))
template <int S>
__device__ __forceinline__ void i8_digits(double xs, int* d) {
    constexpr int W = I8Dig<S>::W;
    unsigned long long bias = 0;
#pragma unroll
    for (int k = 0; k < S - 1; ++k) bias += (1ull << (W - 1)) << (W * k);
    const unsigned long long u = (unsigned long long)__double2ll_rn(xs) + bias;
#pragma unroll
    for (int k = 0; k < S - 1; ++k) d[S - 1 - k] = (int)((u >> (W * k)) & ((1u << W) - 1u)) - (1 << (W - 1));
    d[0] = (int)((long long)u >> (W * (S - 1)));
}

// Digits of FOUR fixed-point values at once, packed one 32-bit word (4 consecutive k) per slice: what the panel builders
// store.  Adding the bias sum_{k < S-1} 64 * 128^k turns the balanced digits into plain bit fields: v + bias =
// sum_k (d_k + 64) 128^k with d_k in [-64, 63] for k < S - 1 and the signed top digit d_{S-1} in [-64, 64] -- an exact
// decomposition of the same integer v as i8_digits' (|digit| <= 64, so the INT32 accumulator bound is unchanged), but the
// fields come out with one shift + mask each and the "- 64" is applied to four packed bytes at once
// (b - 64 in two's complement = flip bit 6, copy the new bit 6 into bit 7).  ~28 integer instructions per value instead
// of ~75 for the residual recursion: the build is bound by the integer pipe as much as by FP64
// (79 integer + 55 FP64 instructions per kernel evaluation before this).
template <int S>
__device__ __forceinline__ void i8_digit_words(const long long (&v)[4], uint32_t (&w)[S]) {
    constexpr int W = I8Dig<S>::W;
    unsigned long long bias = 0;
#pragma unroll
    for (int k = 0; k < S - 1; ++k) bias += (1ull << (W - 1)) << (W * k);
    unsigned long long u[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) u[e] = (unsigned long long)v[e] + bias;
#pragma unroll
    for (int k = 0; k < S - 1; ++k) {
        uint32_t P = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) P |= ((uint32_t)(u[e] >> (W * k)) & ((1u << W) - 1u)) << (8 * e);
        if (W == 8) {
            w[S - 1 - k] = P ^ 0x80808080u;                       // b - 128 in two's complement: flip bit 7
        } else {
            const uint32_t T = P ^ 0x40404040u;                   // b - 64: flip bit 6, copy the new bit 6 into bit 7
            w[S - 1 - k] = T | ((T & 0x40404040u) << 1);
        }
    }
    uint32_t P = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) P |= ((uint32_t)((long long)u[e] >> (W * (S - 1))) & 255u) << (8 * e);
    w[0] = P;
}

// bounded mbarrier wait: a protocol bug must end in a trap (launch failure), never in a hung GPU.  The bound is ~1 minute of
// SM cycles: the longest legitimate wait is one row block of MMAs (milliseconds), but the context may be time-sliced with
// another one on the same GPU (the reference runs its Taichi simulator there) and the cycle counter keeps running meanwhile.
__device__ __forceinline__ void i8_wait(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    for (uint32_t spin = 0;; ++spin) {
        if (mbar_try_wait(bar, parity)) return;
        if ((spin & 1023u) == 1023u && clock64() - t0 > 100000000000LL) {
            printf("sweep_i8_kernel: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// K-major, no swizzle: LBO 128 B (next 16 k), SBO = one 8-row group of the tile = (K bytes / 16) core matrices of 128 B
__device__ __forceinline__ uint64_t i8_desc(uint32_t saddr, uint32_t sbo = (I8_KC / 16) * 128) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// COLL: use of the A-operand collector buffer (SASS A_KEEP / A_REUSE): 0 none, 1 fill (read A from shared memory and keep
// it), 2 use (A is the tile the previous MMA kept), 3 lastuse.  In the sweep's issue order slice s of a stage tile of L^-1
// multiplies S - s panel slices back to back, so A is read from shared memory S times per S (S + 1) / 2 MMAs; without the
// reuse the 128 x 64 x 32 shape is bound by the 6 KB of operand reads per MMA (50 clk), with it by the tensor pipe
// (37 clk, floor 32: profiles/r02_i8_collector_probe.log).
template <int COLL>
__device__ __forceinline__ void i8_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
#define BO_I8_MMA(coll) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8" coll " [%0], %1, %2, %3, p;\n\t}\n" \
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory")
    if (COLL == 1) BO_I8_MMA(".collector::a::fill");
    else if (COLL == 2) BO_I8_MMA(".collector::a::use");
    else if (COLL == 3) BO_I8_MMA(".collector::a::lastuse");
    else BO_I8_MMA("");
#undef BO_I8_MMA
}
__device__ __forceinline__ bool i8_elect() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void i8_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
// exact int32 -> FP64 without the quarter-rate I2F.F64: 2^52 + 2^31 + x is exactly representable, so gluing the biased
// integer under the exponent of 2^52 and subtracting the constant costs one LOP3 and one full-rate DADD
__device__ __forceinline__ double i8_s32_to_f64(int x) {
    return __hiloint2double(0x43300000, x ^ (int)0x80000000) - 4503601774854144.0;
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- operand preparation: row scales and int8 slices of L^-1 in stage-tile order --------------------------------
// dense_cols > 0: the rows are full (B = Ls^T L^-1 of an SVGP state): dense_cols products per row instead of i + 1
__global__ void __launch_bounds__(128) i8_rowscale_kernel(const double* __restrict__ Li, int ld, int np, int n, double* __restrict__ rowscale,
                                                          double* __restrict__ guard_w, int dense_cols = 0) {
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= np) return;
    double m = 0.0;
    const int jend = dense_cols > 0 ? dense_cols : i + 1;
    for (int j = lane; j < jend; j += 32) m = fmax(m, fabs(Li[(size_t)i * ld + j]));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        int e; frexp(m, &e);                          // m = f 2^e, f in [0.5, 1): |x| = |L^-1| 2^-e < 1
        const double rs = (m > 0.0) ? ldexp(1.0, e) : 1.0;
        rowscale[i] = rs;
        // W = max over the real rows of rowscale_i^2 (i + 1): positive doubles order like their bit patterns
        if (i < n) atomicMax(reinterpret_cast<unsigned long long*>(guard_w), (unsigned long long)__double_as_longlong(rs * rs * (double)jend));
    }
}

// linear + Matern kind: guard_w[1] = max_j sum_k lin_w[k] x~_jk^2 (the largest weighted squared norm of a training row), from
// which a candidate's bound on |k*| follows by Cauchy-Schwarz
__global__ void __launch_bounds__(256) i8_xnorm_kernel(const double* __restrict__ Xs, int n, Hyper hyp, double* __restrict__ guard_w) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    double nn = 0.0;
    for (int k = 0; k < hyp.d; ++k) { const double x = Xs[(size_t)j * BO_MAX_DIM + k]; nn = fma(hyp.lin_w[k] * x, x, nn); }
    atomicMax(reinterpret_cast<unsigned long long*>(guard_w + 1), (unsigned long long)__double_as_longlong(nn));
}

// one thread per (row, 16-column chunk) of a stage tile: S x 16 B, tile (ib, kc) at ((ib (ib + 1) / 2) * 2 + kc) * S * 8 KB
// dense = 1 (B = Ls^T L^-1 of an SVGP state): every (ib, kc) tile exists, row-block major, behind the triangular tiles
template <int S>
__global__ void __launch_bounds__(256) i8_pack_linv_kernel(const double* __restrict__ Li, int ld, const double* __restrict__ rowscale,
                                                           int8_t* __restrict__ Lp8, int nbm, int dense = 0) {
    const int ib = blockIdx.y;
    const int kc = blockIdx.x;                        // 0 .. 2 nbm - 1 ; tiles right of the diagonal block do not exist
    if (!dense && kc >= (ib + 1) * (SW_BM / I8_KC)) return;
    const size_t tidx = dense ? (size_t)nbm * (nbm + 1) / 2 * (SW_BM / I8_KC) + (size_t)ib * nbm * (SW_BM / I8_KC) + kc
                              : (size_t)ib * (ib + 1) / 2 * (SW_BM / I8_KC) + kc;
    int8_t* tile = Lp8 + tidx * (size_t)(S * I8_A_SLICE);
    for (int e = threadIdx.x; e < SW_BM * (I8_KC / 16); e += 256) {
        const int r = e / (I8_KC / 16), ch = e % (I8_KC / 16);
        const int i = ib * SW_BM + r, j0 = kc * I8_KC + ch * 16;
        const double inv = ldexp(1.0, I8Dig<S>::F) / rowscale[i];          // exact: both are powers of two
        uint32_t w[S][4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t acc[S];
#pragma unroll
            for (int s = 0; s < S; ++s) acc[s] = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = j0 + q4 * 4 + b;
                const double x = (dense || j <= i) ? Li[(size_t)i * ld + j] * inv : 0.0;
                int dg[S];
                i8_digits<S>(x, dg);
#pragma unroll
                for (int s = 0; s < S; ++s) acc[s] |= (uint32_t)(uint8_t)(int8_t)dg[s] << (8 * b);
            }
#pragma unroll
            for (int s = 0; s < S; ++s) w[s][q4] = acc[s];
        }
        const size_t off = ((size_t)(r / 8) * (I8_KC / 16) + ch) * 128 + (r % 8) * 16;
#pragma unroll
        for (int s = 0; s < S; ++s)
            *reinterpret_cast<uint4*>(tile + (size_t)s * I8_A_SLICE + off) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
}

// ---- the sweep -------------------------------------------------------------------------------------------------
// Warp roles: 0-3 drain the accumulators and run the epilogue, 4 streams stage tiles (TMA), 5 issues the MMAs, 8-11 build
// the NEXT block's int8 panel (FP64 kernel evaluations + slicing) while the tensor core works on the current one.
// Roles are warpgroup-aligned so that setmaxnreg can move registers from the TMA/MMA group (48) to the drain (232: 64 FP64
// column sums + 56 accumulator words per thread) and the builders (208).
template <int DP, int KIND, int S>
__global__ void __launch_bounds__(I8_THREADS, 1) sweep_i8_kernel(const SweepArgs a, const SweepI8Args b) {
    using SM = I8Smem<S, DP>;
    static_assert(S * I8_BN <= 512, "TMEM has 512 columns");
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full  = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    uint64_t* empty = full + I8_STAGES;
    uint64_t* tfull = empty + I8_STAGES;
    uint64_t* tempty = tfull + 1;
    uint64_t* pfull = tempty + 1;             // [2] panel buffer p holds a finished block
    uint64_t* pempty = pfull + 2;             // [2] panel buffer p has been consumed
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(pempty + 2);
    double* colsum  = reinterpret_cast<double*>(smem + SM::OFF_COL);
    double* mu_s    = reinterpret_cast<double*>(smem + SM::OFF_MU);       // [2][64]
    double* tkv     = reinterpret_cast<double*>(smem + SM::OFF_TKV);
    long long* tki  = reinterpret_cast<long long*>(smem + SM::OFF_TKI);
    double* acq_s   = reinterpret_cast<double*>(smem + SM::OFF_ACQ);
    unsigned* cmask = reinterpret_cast<unsigned*>(smem + SM::OFF_CMASK);
    uint32_t* dirs  = reinterpret_cast<uint32_t*>(smem + SM::OFF_SOB);
    uint32_t* shift = dirs + BO_MAX_DIM * BO_SOBOL_BITS;
    double* exp_tab = reinterpret_cast<double*>(smem + SM::OFF_TAB);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int nbm = a.np / SW_BM;
    constexpr int KCH = SW_BM / I8_KC;                        // stages per 128 columns
    constexpr int B_STAGE = S * I8_B_SLICE;
    const size_t panel_bytes = (size_t)(a.np / I8_KC) * B_STAGE;
    int8_t* panel0 = b.panel8 + (size_t)blockIdx.x * 2 * panel_bytes;      // two buffers per CTA

    if (tid == 0) {
        for (int s = 0; s < I8_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1); mbar_init(tempty, 4);
        for (int s = 0; s < 2; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], 1); }
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < BO_MAX_TOPK) { tkv[tid] = -INFINITY; tki[tid] = IDX_EMPTY; }
    if (tid >= 64 && tid < 80) { const double t16[16] = BO_EXP2_16_TABLE; exp_tab[tid - 64] = t16[tid - 64]; }
    if (a.sobol) {
        for (int e = tid; e < DP * BO_SOBOL_BITS; e += I8_THREADS)
            dirs[e] = a.sobol->direction[e / BO_SOBOL_BITS][e % BO_SOBOL_BITS];
        if (tid < DP) shift[tid] = a.sobol->shift[tid];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp >= 8) {
        // ================= builders: candidates, K(X, X*) digits, posterior mean of block it =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;" ::: "memory");
        const int tb = tid - 256, wb = warp - 8;
        int it = 0;
        for (long long blk = blockIdx.x; blk < a.nblocks; blk += gridDim.x, ++it) {
            const int p = it & 1;
            i8_wait(&pempty[p], ((it >> 1) & 1) ^ 1);         // the block that used this buffer two turns ago is done
            int8_t* panel = panel0 + (size_t)p * panel_bytes;
            double xc[2][DP];
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                long long li = blk * I8_BN + (wb + 4 * gi) * 8 + g;
                if (li >= a.N) li = a.N - 1;
                if (a.cand) {
#pragma unroll
                    for (int k = 0; k < DP; ++k) xc[gi][k] = (k < a.d) ? a.cand[(size_t)li * a.d + k] : 0.0;
                } else {
                    sobol_point<DP>(dirs, shift, a.d, a.first_index + li, xc[gi]);
                }
#pragma unroll
                for (int k = 0; k < DP; ++k) xc[gi][k] = __dmul_rn(xc[gi][k], a.hyp.inv_ls[k]);     // (never contracted into the differences below: every code instance must round alike)
            }
            double mu0 = 0.0, mu1 = 0.0;
            // X~ and alpha staged through the builders' own shared-memory region in double-buffered cp.async chunks
            constexpr int XCH = SM::XCH, XP = DP + 2, XBUF = XCH * XP + XCH;
            double* xstage = reinterpret_cast<double*>(smem + SM::OFF_X);
            const int nrows = a.np;
            const int nchunks = (nrows + XCH - 1) / XCH;
            auto load_chunk = [&](int c) {
                double* xb = xstage + (c & 1) * XBUF;
                double* ab = xb + XCH * XP;
                const int r0 = c * XCH, rows = min(XCH, nrows - r0);
                for (int e = tb; e < rows * (DP / 2); e += 128) {
                    const int r = e / (DP / 2), k = e % (DP / 2);
                    // rows are staged permuted (slot (r % 4) * (XCH / 4) + r / 4): the rows 4 q + e that the four q-lanes of a
                    // candidate read together become adjacent slots -> conflict-free LDS.128 (natural order: 2-way conflicts)
                    cp_async16(xb + ((r & 3) * (XCH / 4) + (r >> 2)) * XP + 2 * k, a.Xs + (size_t)(r0 + r) * BO_MAX_DIM + 2 * k);
                }
                for (int e = tb; e < rows / 2; e += 128) cp_async16(ab + 2 * e, a.alpha + r0 + 2 * e);
                cp_async_commit();
            };
            load_chunk(0);
            for (int c = 0; c < nchunks; ++c) {
                if (c + 1 < nchunks) { load_chunk(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const double* xb = xstage + (c & 1) * XBUF;
                const double* ab = xb + XCH * XP;
                const int jend = min(nrows, (c + 1) * XCH);
                // 32 rows per iteration: this lane evaluates rows j0 + 16 h + 4 q + e (h < 2, e < 4) for its two candidates,
                // i.e. one 32-bit word (4 consecutive k) of every slice per (candidate, h)
                for (int j0 = c * XCH; j0 < jend; j0 += 32) {
                    double kv[2][8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const int j = j0 + 16 * (r >> 2) + 4 * q + (r & 3);
                        const int jr = j - c * XCH;
                        double x[DP];
                        const double2* row = reinterpret_cast<const double2*>(xb + ((jr & 3) * (XCH / 4) + (jr >> 2)) * XP);
#pragma unroll
                        for (int k = 0; k < DP / 2; ++k) { const double2 t = row[k]; x[2 * k] = t.x; x[2 * k + 1] = t.y; }
                        const double al = ab[jr];
#pragma unroll
                        for (int gi = 0; gi < 2; ++gi) {
                            double sq = 0.0;
#pragma unroll
                            for (int k = 0; k < DP; ++k) { const double df = __dsub_rn(xc[gi][k], x[k]); sq = fma(df, df, sq); }
                            // (the lean square root / table exponential of the CTA-pair kernel's build: this kernel is build-bound at the
                            // small n it serves -- four builder warps against 20 stages of MMAs per block at n = 512)
                            const double v = kernel_value_fast_t<KIND>(sq, a.hyp.outputscale, exp_tab);
                            kv[gi][r] = (j < a.n) ? v : 0.0;
                        }
                        mu0 = fma(kv[0][r], al, mu0); mu1 = fma(kv[1][r], al, mu1);
                    }
                    int8_t* st_tile = panel + (size_t)(j0 / I8_KC) * B_STAGE;
                    const int ch0 = (j0 % I8_KC) / 16;
#pragma unroll
                    for (int gi = 0; gi < 2; ++gi)
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            uint32_t w[S];
                            long long v4[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) v4[e] = __double2ll_rn(kv[gi][hh * 4 + e] * b.dig_scale);
                            i8_digit_words<S>(v4, w);
                            const size_t off = ((size_t)(wb + 4 * gi) * (I8_KC / 16) + ch0 + hh) * 128 + g * 16 + q * 4;
#pragma unroll
                            for (int s = 0; s < S; ++s) *reinterpret_cast<uint32_t*>(st_tile + (size_t)s * I8_B_SLICE + off) = w[s];
                        }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");     // buffer free before chunk c + 2 overwrites it
            }
            mu0 += __shfl_xor_sync(0xffffffffu, mu0, 1); mu0 += __shfl_xor_sync(0xffffffffu, mu0, 2);
            mu1 += __shfl_xor_sync(0xffffffffu, mu1, 1); mu1 += __shfl_xor_sync(0xffffffffu, mu1, 2);
            if (q == 0) { mu_s[p * I8_BN + wb * 8 + g] = mu0; mu_s[p * I8_BN + (wb + 4) * 8 + g] = mu1; }
            __threadfence();
            fence_proxy_async();      // generic-proxy panel writes -> visible to the async-proxy (TMA) reads
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tb == 0) mbar_arrive(&pfull[p]);
        }
    } else if (warp >= 4) {
        // ================= warpgroup 1: TMA producer (warp 4), MMA issuer (warp 5) =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;" ::: "memory");
        if (warp < 6) {
        // flags bit 2: clock64 accounting of where the MMA issuer waits (printed by CTA 0 at the end; triage only)
        const bool prof = (a.flags & 4) != 0;
        long long t_tot = 0, t_w0 = 0, t_w1 = 0, t_w2 = 0;
        unsigned long long ns0 = 0, ns1 = 0;
        if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
        const long long c0 = prof ? clock64() : 0;
        int stage = 0; uint32_t phase = 0;        // ring position of the role this thread plays (producer or MMA issuer)
        uint32_t rb = 0;                          // running row-block counter (accumulator full/empty phases)
        int it = 0;
        for (long long blk = blockIdx.x; blk < a.nblocks; blk += gridDim.x, ++it) {
            const int p = it & 1;
            const int8_t* panel = panel0 + (size_t)p * panel_bytes;
            const long long tB0 = prof ? clock64() : 0;
            i8_wait(&pfull[p], (it >> 1) & 1);            // the builders have finished this block (acquire: panel)
            if (prof) t_w2 += clock64() - tB0;
            if (warp == 4) {
                if (lane == 0) {
                    // (an L2 prefetch of the tiles four stages ahead was tried: the issuer's stage-full waits fell from 34.2 M to
                    // 31.7 M cycles per 254 blocks, but the builders slowed down by more -- the ring is bound by shared-memory
                    // bandwidth, MMA operand reads + TMA writes, not by HBM latency)
                    for (int ib = 0; ib < nbm; ++ib)
                        for (int kc = 0; kc < (ib + 1) * KCH; ++kc) {
                            i8_wait(&empty[stage], phase ^ 1);
                            unsigned char* sb = smem + stage * SM::STAGE_BYTES;
                            mbar_expect_tx(&full[stage], SM::STAGE_BYTES);
                            bulk_g2s(sb, b.Lp8 + ((size_t)ib * (ib + 1) / 2 * KCH + kc) * (size_t)(S * I8_A_SLICE), S * I8_A_SLICE, &full[stage]);
                            bulk_g2s(sb + S * I8_A_SLICE, panel + (size_t)kc * B_STAGE, B_STAGE, &full[stage]);
                            if (++stage == I8_STAGES) { stage = 0; phase ^= 1; }
                        }
                }
            } else {
                // The whole warp walks the loop (warp-uniform control flow keeps descriptors in uniform registers); one elected
                // lane issues.  The slice-pair loops are fully unrolled: every descriptor is the stage's base descriptor plus a
                // compile-time constant, so an MMA costs a few integer instructions to issue -- with rolled loops the issuing
                // thread, not the tensor core, set the pace (a 64-column MMA lasts ~34 cycles).
                const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_BN >> 3) << 17) | ((uint32_t)(SW_BM >> 4) << 24);
                const bool leader = i8_elect();
                for (int ib = 0; ib < nbm; ++ib, ++rb) {
                    const long long w1 = prof ? clock64() : 0;
                    i8_wait(tempty, (rb & 1) ^ 1);             // the drain of the previous row block is done
                    if (prof) t_w1 += clock64() - w1;
                    tc_fence_after();
                    const int nkc = (ib + 1) * KCH;
                    for (int kc = 0; kc < nkc; ++kc) {
                        const long long w0 = prof ? clock64() : 0;
                        i8_wait(&full[stage], phase);
                        if (prof) t_w0 += clock64() - w0;
                        tc_fence_after();
                        const uint32_t a0 = smem_u32(smem + stage * SM::STAGE_BYTES);
                        const uint64_t da0 = i8_desc(a0), db0 = i8_desc(a0 + S * I8_A_SLICE);
                        const uint32_t acc0 = kc > 0 ? 1u : 0u;
                        if (leader) {
#pragma unroll
                            for (int kk = 0; kk < I8_KC / 32; ++kk)
#pragma unroll
                                for (int s = 0; s < S; ++s)
#pragma unroll
                                    for (int t = 0; t + s < S; ++t) {
                                        const uint32_t td = tmem_base + (s + t) * I8_BN;
                                        const uint64_t da = da0 + (uint64_t)((s * I8_A_SLICE + kk * 256) >> 4);
                                        const uint64_t db = db0 + (uint64_t)((t * I8_B_SLICE + kk * 256) >> 4);
                                        const uint32_t accf = (kk > 0 || s > 0) ? 1u : acc0;
                                        if (s == S - 1)          i8_mma<0>(td, da, db, idesc, accf);     // a run of one
                                        else if (t == 0)         i8_mma<1>(td, da, db, idesc, accf);
                                        else if (t == S - 1 - s) i8_mma<3>(td, da, db, idesc, accf);
                                        else                     i8_mma<2>(td, da, db, idesc, accf);
                                    }
                            i8_commit(&empty[stage]);                // the slot is free once these MMAs have read it
                            if (kc == nkc - 1) i8_commit(tfull);     // the accumulators of row block ib are complete
                        }
                        __syncwarp();
                        if (++stage == I8_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
            asm volatile("bar.sync 2, 192;" ::: "memory");
            if (prof) t_tot += clock64() - tB0;
            asm volatile("bar.sync 2, 192;" ::: "memory");
            asm volatile("bar.sync 2, 192;" ::: "memory");
        }
        if (prof && blockIdx.x == 0 && tid == 160) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
            const long long c1 = clock64();
            printf("sweep_i8 CTA 0 MMA issuer: %d blocks, %lld clk in the contraction; waits: panel-ready %lld, stage-full %lld, accumulators-drained %lld; "
                   "SM clock over the kernel %.0f MHz\n", it, t_tot, t_w2, t_w0, t_w1, (double)(c1 - c0) / (double)(ns1 - ns0) * 1e3);
        }
        }
    } else {
        // ================= warpgroup 0: accumulator drain + epilogue =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
        uint32_t rb = 0;
        int it = 0;
        for (long long blk = blockIdx.x; blk < a.nblocks; blk += gridDim.x, ++it) {
            const int p = it & 1;
            i8_wait(&pfull[p], (it >> 1) & 1);            // acquire: mu_s[p]
            {
                double acc[I8_BN];
#pragma unroll
                for (int c = 0; c < I8_BN; ++c) acc[c] = 0.0;
                for (int ib = 0; ib < nbm; ++ib, ++rb) {
                    const double rs = b.rowscale[ib * SW_BM + tid] * b.eb_scale;
                    i8_wait(tfull, rb & 1);
                    tc_fence_after();
                    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
                    // the accumulators go back to the MMA issuer as soon as the last chunk has landed in registers, before its
                    // arithmetic (a software-pipelined variant with two x4 register sets measured slower: 50.5 M vs 45.4 M cycles of
                    // issuer wait per 254 blocks, profiles/r02_i8_role_wait_accounting.log)
#pragma unroll
                    for (int c0 = 0; c0 < I8_BN; c0 += 8) {
                        int v[S][8];
#pragma unroll
                        for (int gq = 0; gq < S; ++gq) tmem_ld8(trow + gq * I8_BN + c0, v[gq]);
                        tmem_ld_wait();
#pragma unroll
                        for (int gq = 0; gq < S; ++gq)      // (tcgen05.ld writes asynchronously: no use may be scheduled above the wait)
                            asm volatile("" : "+r"(v[gq][0]), "+r"(v[gq][1]), "+r"(v[gq][2]), "+r"(v[gq][3]), "+r"(v[gq][4]), "+r"(v[gq][5]), "+r"(v[gq][6]), "+r"(v[gq][7]));
                        if (c0 + 8 == I8_BN) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(tempty);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            double t = i8_s32_to_f64(v[S - 1][j]);
#pragma unroll
                            for (int gq = S - 2; gq >= 0; --gq) t = fma(t, I8Dig<S>::HORNER, i8_s32_to_f64(v[gq][j]));
                            const double u = t * rs;
                            acc[c0 + j] = fma(u, u, acc[c0 + j]);
                        }
                    }
                }
                // sum over the 32 rows of this warp: halving butterfly (each step trades half of the columns held)
#pragma unroll
                for (int o = 16, cnt = I8_BN / 2; o >= 1; o >>= 1, cnt >>= 1) {
                    const bool upper = (lane & o) != 0;
#pragma unroll
                    for (int c = 0; c < cnt; ++c) {
                        const double send = upper ? acc[c] : acc[c + cnt];
                        const double keep = upper ? acc[c + cnt] : acc[c];
                        acc[c] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                    }
                }
                // lane l now holds the warp totals of 2 columns; recover which ones from the butterfly's bit order
                {
                    int base = 0;
#pragma unroll
                    for (int o = 16, cnt = I8_BN / 2; o >= 1; o >>= 1, cnt >>= 1) base += (lane & o) ? cnt : 0;
                    colsum[warp * I8_BN + base] = acc[0];
                    colsum[warp * I8_BN + base + 1] = acc[1];
                }
            }
            asm volatile("bar.sync 2, 192;" ::: "memory");

            // ================= epilogue: variance, acquisition, CTA-local top-k =====================
            if (tid < I8_BN) {
                const long long li = blk * I8_BN + tid;
                const double ss = ((colsum[tid] + colsum[I8_BN + tid]) + (colsum[2 * I8_BN + tid] + colsum[3 * I8_BN + tid])) * b.ss_scale;
                // accuracy guard: too little variance left for the slicing error bound -> the FP64 contraction scores it
                const bool flagged = b.flag_count != nullptr && li < a.N &&
                                     !(a.hyp.outputscale - ss >= b.guard_scale * sqrt(__ldg(b.guard_w) * ss));
                {
                    const unsigned fm = __ballot_sync(0xffffffffu, flagged);
                    if (fm) {
                        int base = 0;
                        if (lane == __ffs(fm) - 1) base = atomicAdd(b.flag_count, __popc(fm));
                        base = __shfl_sync(0xffffffffu, base, __ffs(fm) - 1);
                        const long long pos = base + __popc(fm & ((1u << lane) - 1u));
                        if (flagged && pos < b.flag_cap) b.flag_idx[pos] = li;
                    }
                }
                const double var = fmax(a.hyp.outputscale - ss, a.min_var);
                const double mean = a.hyp.mean + mu_s[p * I8_BN + tid];
                double v = acq_value(a.acq, mean, var, a.best_f, a.sqrt_beta);
                if (li < a.N) {
                    if (a.mean_out) a.mean_out[li] = mean;
                    if (a.var_out) a.var_out[li] = var;
                    if (a.acq_out) a.acq_out[li] = v;
                }
                if (!(v == v)) v = -INFINITY;
                acq_s[tid] = v;
                bool beats = false;
                if (a.topk > 0 && li < a.N && !flagged) beats = tk_better(v, a.first_index + li, tkv[a.topk - 1], tki[a.topk - 1]);
                const unsigned m = __ballot_sync(0xffffffffu, beats);
                if (lane == 0) cmask[warp] = m;
            }
            asm volatile("bar.sync 2, 192;" ::: "memory");
            if (tid == 0) {
                mbar_arrive(&pempty[p]);           // panel p and mu_s[p] are free for the builders (block it + 2)
                if (a.topk > 0) {
                    const int K = a.topk;
                    for (int w = 0; w < I8_BN / 32; ++w) {
                        unsigned m = cmask[w];
                        while (m) {
                            const int c = w * 32 + __ffs(m) - 1;
                            m &= m - 1;
                            const double v = acq_s[c];
                            const long long gi = a.first_index + blk * I8_BN + c;
                            if (!tk_better(v, gi, tkv[K - 1], tki[K - 1])) continue;
                            int pp = K - 1;
                            while (pp > 0 && tk_better(v, gi, tkv[pp - 1], tki[pp - 1])) { tkv[pp] = tkv[pp - 1]; tki[pp] = tki[pp - 1]; --pp; }
                            tkv[pp] = v; tki[pp] = gi;
                        }
                    }
                }
            }
            asm volatile("bar.sync 2, 192;" ::: "memory");
        }
        if (tid < BO_MAX_TOPK && a.part_val) {
            a.part_val[(size_t)blockIdx.x * BO_MAX_TOPK + tid] = tkv[tid];
            a.part_idx[(size_t)blockIdx.x * BO_MAX_TOPK + tid] = tki[tid];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

#include "sweep_i8_pair.cuh"

// ---- host side -------------------------------------------------------------------------------------------------
// the model side of eligibility: exact GP with a stationary kernel (|k*| <= output scale), more than one stage of rows
static bool sweep_i8_model_ok(const bo_handle* h) {
    // np < 2^16: the INT32 group accumulators hold at most 8 * 64 * 64 * np
    // (the linear + Matern kind runs on the CTA-pair kernel only: per-candidate operand scale)
    // (so does an SVGP predictive state: its stacked factor [L^-1; Ls^T L^-1] exists in the pair kernel only)
    return h->fitted && (!h->svgp || (h->svB != nullptr && h->sm_count >= 2)) && h->np >= I8_MIN_NP && h->np < 65536;
}

// Slice count of AUTO.  Both sliced forms carry the per-candidate accuracy guard, so both deliver the same guarantee
// (slicing error of an unflagged candidate <= 2e-9 relative); they differ in how many candidates the guard sends to the FP64
// pass and in cost.  Seven slices with 8-bit lower digits (54-bit operands, 28 products) have a 4x larger bound than eight
// 7-bit slices (55 bits, 36 products): on the reference's CSV rows that is a few hundred more re-scored candidates out of
// 20 000, on pools that stay off the data none -- for 22 % fewer MMAs on a power-bound kernel.  AUTO takes the 7-slice form
// wherever its INT32 accumulators cannot overflow (np <= 16 384) and the 8-slice form above.
static int sweep_i8_slices(const bo_handle* h) { return h->np <= I8Dig<7>::NP_MAX ? 7 : 8; }

// The pinned mode a sweep over a pool of `pool` candidates runs in.  Pinned modes depend on the model only, so every
// shard of a pool takes the same path and the per-candidate values are bit-identical for every shard layout; AUTO also
// looks at the pool size (small pools keep the row-split FP64 kernel, which fills the SMs with fewer than two waves of
// 64-candidate blocks) -- callers that shard a pool resolve AUTO once on the global size (bo_resolve_sweep_mode).
int resolve_sweep_mode(const bo_handle* h, int mode, long long pool) {
    if (mode == BO_SWEEP_FP64 || !sweep_i8_model_ok(h)) return BO_SWEEP_FP64;
    if (mode == BO_SWEEP_I8X7) return h->np <= I8Dig<7>::NP_MAX ? BO_SWEEP_I8X7 : BO_SWEEP_I8X8;      // accumulator bound
    if (mode == BO_SWEEP_I8X8) return mode;
    // (an SVGP state has no row-split FP64 kernel to fall back on: one full turn of the CTA pairs is enough there)
    if ((pool + I8_BN - 1) / I8_BN < (h->svgp ? h->sm_count / 2 : 2LL * h->sm_count)) return BO_SWEEP_FP64;
    if (h->np < I8_AUTO_MIN_NP) return BO_SWEEP_FP64;     // measured gain starts at n = 512 (1.33x); below it was not measured
    // no hyper-parameter heuristic: every sliced sweep carries the per-candidate accuracy guard (I8_GUARD_*), which sends
    // the candidates the slicing error could matter for -- sigma^2 orders of magnitude below the prior variance, next to
    // training rows -- through the FP64 contraction
    return sweep_i8_slices(h) == 7 ? BO_SWEEP_I8X7 : BO_SWEEP_I8X8;
}

template <int S>
static int ensure_i8_ws(bo_handle* h, int grid, long long pool) {
    const int nbm = h->np / SW_BM;
    // packed tiles + one all-zero stage tile behind them (Lp8_zero)
    const size_t a_bytes = ((size_t)nbm * (nbm + 1) / 2 * (SW_BM / I8_KC) + (h->svgp ? (size_t)nbm * nbm * (SW_BM / I8_KC) : 0) + 1) *
                           (size_t)(S * I8_A_SLICE);
    if (a_bytes > h->Lp8_bytes) {
        if (h->Lp8) cudaFree(h->Lp8);
        h->Lp8 = nullptr; h->Lp8_bytes = 0; h->Lp8_epoch = 0;
        BO_CUDA(h, cudaMalloc(&h->Lp8, a_bytes));
        h->Lp8_bytes = a_bytes;
    }
    if (2 * (size_t)h->np > h->rowscale_cap) {                 // (second half: the rows of B of an SVGP state)
        if (h->rowscale) cudaFree(h->rowscale);
        h->rowscale = nullptr; h->rowscale_cap = 0; h->Lp8_epoch = 0;
        BO_CUDA(h, cudaMalloc(&h->rowscale, 2 * (size_t)h->cap_np * sizeof(double)));
        h->rowscale_cap = 2 * (size_t)h->cap_np;
    }
    const size_t p_bytes = (size_t)grid * 2 * (h->np / I8_KC) * (size_t)(S * I8_B_SLICE);      // two panel buffers per CTA
    if (p_bytes > h->panel8_bytes) {
        if (h->panel8) cudaFree(h->panel8);
        h->panel8 = nullptr; h->panel8_bytes = 0;
        BO_CUDA(h, cudaMalloc(&h->panel8, p_bytes));
        h->panel8_bytes = p_bytes;
    }
    // guard: room for every candidate of the pool (8 B each).  The path a candidate takes must depend on the candidate alone --
    // a "too many flagged, re-score the whole pool" shortcut would make values depend on the shard layout
    if (!h->guard_dev) { BO_CUDA(h, cudaMalloc(&h->guard_dev, 4 * sizeof(double))); h->Lp8_epoch = 0; }
    if (!h->flag_count_dev) BO_CUDA(h, cudaMalloc(&h->flag_count_dev, sizeof(int)));
    if (!h->flag_count_host) BO_CUDA(h, cudaMallocHost(&h->flag_count_host, sizeof(int)));
    const size_t cap = (size_t)(pool > 0 ? pool : 1);
    if (cap > h->flag_cap) {
        if (h->flag_idx) cudaFree(h->flag_idx);
        h->flag_idx = nullptr; h->flag_cap = 0;
        BO_CUDA(h, cudaMalloc(&h->flag_idx, cap * sizeof(long long)));
        h->flag_cap = cap;
    }
    return 0;
}

template <int DP, int KIND, int S>
static int launch_sweep_i8_k(bo_handle* h, const SweepArgs& a, const SweepI8Args& b, int grid, cudaStream_t st) {
    BO_CUDA(h, cudaFuncSetAttribute(sweep_i8_kernel<DP, KIND, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8Smem<S, DP>::BYTES));
    sweep_i8_kernel<DP, KIND, S><<<grid, I8_THREADS, I8Smem<S, DP>::BYTES, st>>>(a, b);
    BO_LAUNCH_CHECK(h);
    return 0;
}
template <int DP>
static int launch_sweep_i8(bo_handle* h, const SweepArgs& a, const SweepI8Args& b, int S, int grid, cudaStream_t st) {
    if (a.hyp.kind == BO_KERNEL_MATERN52)
        return S == 8 ? launch_sweep_i8_k<DP, BO_KERNEL_MATERN52, 8>(h, a, b, grid, st) : launch_sweep_i8_k<DP, BO_KERNEL_MATERN52, 7>(h, a, b, grid, st);
    return S == 8 ? launch_sweep_i8_k<DP, BO_KERNEL_RBF, 8>(h, a, b, grid, st) : launch_sweep_i8_k<DP, BO_KERNEL_RBF, 7>(h, a, b, grid, st);
}

// a: as prepared by sweep_impl.  Slices the current factor if it changed since the last sliced sweep, sweeps, then re-scores
// the candidates the accuracy guard flagged on the FP64 contraction (same pool positions, same outputs, same top-k order).
// Synchronises the stream once (the 4-byte count of flagged candidates decides whether a second pass is needed).
static int sweep_i8_run(bo_handle* h, const SweepArgs& a_in, int S, double* vals_dev, int64_t* idx_dev, cudaStream_t st) {
    SweepArgs a = a_in;
    // the L2 eviction hints pin ONE model's L^-1 tiles; an SVGP scan runs T handles' stacked factors side by side (8 x 44 MB at
    // M = 2048), where "evict last" for all of them measured 2-3 % slower than no hint (profiles/r02_l2_hint_ab.log)
    if (h->svgp && !getenv("BO_B200_SWEEP_FLAGS")) a.flags &= ~(8 | 16 | 32 | 64);
    a.G = 1; a.seg[0] = 0; a.seg[1] = h->np / SW_BM;
    a.nblocks = (a.N + I8_BN - 1) / I8_BN;
    // CTA pairs (cta_group::2, sweep_i8_pair.cuh) share one 64-candidate block; BO_B200_I8_PAIR=0 keeps the one-CTA kernel (triage / A-B)
    // The pair kernel builds a block's panel BEFORE its MMAs (FP64 crawls under the INT8 tensor pipe), which pays once the MMA
    // phase dominates; small models (few row blocks per candidate block) keep the one-CTA kernel, whose build overlaps the
    // MMAs: n = 512: 121 M cand/s one-CTA vs 74 M pair; n = 1024: 49 vs 35; n = 4096: 3.8 vs 4.1 (profiles/r02_*).
    const char* pe = getenv("BO_B200_I8_PAIR");
    const bool pairs = h->sm_count >= 2 && (pe ? atoi(pe) != 0 : (h->np >= I8_PAIR_MIN_NP || h->hyp.kind == BO_KERNEL_LINEAR_MATERN52 || h->svgp));
    if (!pairs && (h->hyp.kind == BO_KERNEL_LINEAR_MATERN52 || h->svgp))   // the one-CTA kernel has no per-candidate operand scale / stacked factor
        return sweep_fp64_run(h, a_in, 0, true, vals_dev, idx_dev, nullptr, st);
    int grid;
    if (pairs) {
        const long long np2 = a.nblocks < h->sm_count / 2 ? a.nblocks : h->sm_count / 2;
        grid = 2 * (int)np2;
    } else {
        grid = (int)(a.nblocks < h->sm_count ? a.nblocks : h->sm_count);
    }
    { const char* gs = getenv("BO_B200_I8_GRID"); if (gs && atoi(gs) >= 2 && atoi(gs) < grid) grid = atoi(gs) & ~1; }   // triage: fewer CTAs
    int rc;
    if ((rc = ensure_sweep_ws(h, grid, false, grid + h->sm_count))) return rc;    // top-k lists: this kernel's + the re-score pass's
    if ((rc = (S == 8 ? ensure_i8_ws<8>(h, grid, a.N) : ensure_i8_ws<7>(h, grid, a.N)))) return rc;
    const int nbm = h->np / SW_BM;
    const size_t tiles_all = (size_t)nbm * (nbm + 1) / 2 * (SW_BM / I8_KC) + (h->svgp ? (size_t)nbm * nbm * (SW_BM / I8_KC) : 0);
    if (h->Lp8_epoch != h->factor_epoch || h->Lp8_S != S) {
        // the factor changed since the operands were sliced (fit, append, refit) or the slice count did
        BO_CUDA(h, cudaMemsetAsync(h->guard_dev, 0, 4 * sizeof(double), st));
        i8_rowscale_kernel<<<(h->np + 3) / 4, 128, 0, st>>>(h->Li, h->cap_np, h->np, h->n, h->rowscale, h->guard_dev);
        BO_LAUNCH_CHECK(h);
        if (S == 8) i8_pack_linv_kernel<8><<<dim3(nbm * (SW_BM / I8_KC), nbm), 256, 0, st>>>(h->Li, h->cap_np, h->rowscale, h->Lp8, nbm);
        else        i8_pack_linv_kernel<7><<<dim3(nbm * (SW_BM / I8_KC), nbm), 256, 0, st>>>(h->Li, h->cap_np, h->rowscale, h->Lp8, nbm);
        BO_LAUNCH_CHECK(h);
        if (h->hyp.kind == BO_KERNEL_LINEAR_MATERN52) {
            i8_xnorm_kernel<<<(h->n + 255) / 256, 256, 0, st>>>(h->Xs, h->n, h->hyp, h->guard_dev);
            BO_LAUNCH_CHECK(h);
        }
        if (h->svgp) {          // the rows of B = Ls^T L^-1 behind those of L^-1: their scales, their W (guard_w[2]), their dense tiles
            i8_rowscale_kernel<<<(h->np + 3) / 4, 128, 0, st>>>(h->svB, h->np, h->np, h->n, h->rowscale + h->np, h->guard_dev + 2, h->n);
            BO_LAUNCH_CHECK(h);
            if (S == 8) i8_pack_linv_kernel<8><<<dim3(nbm * (SW_BM / I8_KC), nbm), 256, 0, st>>>(h->svB, h->np, h->rowscale + h->np, h->Lp8, nbm, 1);
            else        i8_pack_linv_kernel<7><<<dim3(nbm * (SW_BM / I8_KC), nbm), 256, 0, st>>>(h->svB, h->np, h->rowscale + h->np, h->Lp8, nbm, 1);
            BO_LAUNCH_CHECK(h);
        }
        BO_CUDA(h, cudaMemsetAsync(h->Lp8 + tiles_all * (size_t)(S * I8_A_SLICE), 0, (size_t)S * I8_A_SLICE, st));
        h->Lp8_epoch = h->factor_epoch; h->Lp8_S = S;
    }
    SweepI8Args b{};
    b.Lp8 = h->Lp8; b.rowscale = h->rowscale; b.panel8 = h->panel8;
    b.Lp8_zero = h->Lp8 + tiles_all * (size_t)(S * I8_A_SLICE);
    b.sv = h->svgp ? 1 : 0; b.sv_add = h->svgp ? h->sv_add : 0.0;
    int e; frexp(a.hyp.outputscale, &e);                      // |k*| <= outputscale < 2^e
    b.dig_scale = ldexp(1.0, (S == 7 ? I8Dig<7>::F : I8Dig<8>::F) - e);
    b.eb_scale = ldexp(1.0, -12);
    b.ss_scale = ldexp(1.0, 2 * e);
    const char* ng = getenv("BO_B200_I8_NO_GUARD");           // triage only: raw sliced values for every candidate
    const bool guard = !(ng && atoi(ng) == 1);
    // (linear + Matern kind: the pair kernel multiplies by the candidate's own eb instead of 2^e)
    b.guard_scale = I8_GUARD_KAPPA / I8_GUARD_RTOL *
                    ldexp(1.0, (h->hyp.kind == BO_KERNEL_LINEAR_MATERN52 ? 0 : e) + (S == 7 ? I8Dig<7>::GEXP : I8Dig<8>::GEXP));
    b.guard_w = h->guard_dev;
    b.flag_idx = h->flag_idx; b.flag_count = guard ? h->flag_count_dev : nullptr; b.flag_cap = (long long)h->flag_cap;
    a.part_val = h->part_val; a.part_idx = (long long*)h->part_idx;
    BO_CUDA(h, cudaMemsetAsync(h->flag_count_dev, 0, sizeof(int), st));
    BO_CUDA(h, cudaEventRecord(h->ev0, st));
    if ((rc = pairs ? BO_DISPATCH_DP(h->dp, launch_sweep_i8_pair, h, a, b, S, grid, st)
                    : BO_DISPATCH_DP(h->dp, launch_sweep_i8, h, a, b, S, grid, st))) return rc;
    BO_CUDA(h, cudaEventRecord(h->ev1, st));
    h->sweep_timed = true;
    BO_CUDA(h, cudaMemcpyAsync(h->flag_count_host, h->flag_count_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    BO_CUDA(h, cudaStreamSynchronize(st));
    const long long flagged = *h->flag_count_host;
    int lists = grid;
    if (flagged > 0) {
        SweepArgs a2 = a_in;
        a2.N = flagged; a2.idx_map = h->flag_idx;
        int g2 = 0;
        if ((rc = sweep_fp64_run(h, a2, grid, false, nullptr, nullptr, &g2, st))) return rc;
        lists += g2;
    }
    h->sweep_path = S; h->sweep_flagged = flagged;
    if (a.topk > 0) {
        topk_merge_kernel<<<1, 1024, 0, st>>>(h->part_val, (long long*)h->part_idx, lists * BO_MAX_TOPK, a.topk, vals_dev, (long long*)idx_dev);
        BO_LAUNCH_CHECK(h);
    }
    return 0;
}

// ---- INT8 tensor-pipe peak probe (roofline denominator for the sliced sweep): 128 x 256 x 32 MMAs on resident operands ----
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters) {
    constexpr int N = 256, KC = 64;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (SW_BM + N) * KC / 4; i += 128) {
        uint32_t hsh = (uint32_t)i * 0x9E3779B1u; hsh ^= hsh >> 15; hsh *= 0x2C1B3C6Du; hsh ^= hsh >> 12;
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) w |= (uint32_t)(uint8_t)(int8_t)((int)((hsh >> (8 * b)) & 127u) - 64) << (8 * b);
        reinterpret_cast<uint32_t*>(smem)[i] = w;
    }
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(SW_BM >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + SW_BM * KC;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int kk = 0; kk < KC / 32; ++kk)
                    i8_mma<0>(tmem_base + (r & 1) * N, i8_desc(a0 + kk * 256, 512), i8_desc(b0 + kk * 256, 512), idesc, (it | r | kk) ? 1u : 0u);
            i8_commit(&bars[it & 1]);
            if (it > 0) i8_wait(&bars[(it - 1) & 1], ((it - 1) >> 1) & 1);
        }
        i8_wait(&bars[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

int i8_peak_impl(bo_handle* h, double seconds, double* tops) {
    BO_CUDA(h, cudaSetDevice(h->device));
    const int smem_bytes = (SW_BM + 256) * 64, iters = 4000;
    BO_CUDA(h, cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const double ops_per_launch = (double)h->sm_count * iters * 16.0 * SW_BM * 256 * 32 * 2.0;
    cudaEvent_t e0, e1;
    BO_CUDA(h, cudaEventCreate(&e0));
    BO_CUDA(h, cudaEventCreate(&e1));
    auto launch = [&]() { i8_peak_kernel<<<h->sm_count, 128, smem_bytes>>>(iters); h->launches++; };
    launch();
    BO_CUDA(h, cudaDeviceSynchronize());
    BO_CUDA(h, cudaEventRecord(e0));
    launch();
    BO_CUDA(h, cudaEventRecord(e1));
    BO_CUDA(h, cudaEventSynchronize(e1));
    float ms1 = 0.f;
    BO_CUDA(h, cudaEventElapsedTime(&ms1, e0, e1));
    int reps = (int)(seconds * 1e3 / (ms1 > 0.01f ? ms1 : 0.01f));
    reps = reps < 1 ? 1 : reps > 2000 ? 2000 : reps;
    BO_CUDA(h, cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    BO_CUDA(h, cudaEventRecord(e1));
    BO_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0.f;
    BO_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    BO_CUDA(h, cudaGetLastError());
    *tops = ops_per_launch * reps / (ms * 1e-3) * 1e-12;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
