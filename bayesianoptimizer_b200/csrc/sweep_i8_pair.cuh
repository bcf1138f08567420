// K4-i8, CTA-pair form: the sliced sweep of sweep_i8.cuh with tcgen05.mma cta_group::2 -- included by sweep.cu after it.
//
// Why: in the one-CTA kernel the 128 x 64 x 32 MMA is fed from shared memory that the TMA ring is refilling at the same
// time: 2.9 KB of operand reads (with the A-collector reuse) + 1.33 KB of TMA writes per MMA against 128 B/clk leave the
// tensor pipe waiting (41.5 clk per MMA in the kernel, 36.7 in isolation, floor 32; stage-full and drain waits on top;
// profiles/r02_i8_role_wait_accounting.log).  A CTA pair shares one 64-candidate block: the MMA is M = 256 -- CTA r of
// the pair holds row block 2 i + r of L^-1 (its own 128 TMEM lanes, its own drain) and HALF of the panel (32 candidates,
// which it builds itself) -- so per SM the B reads and the B copies halve (3.0 KB per MMA: math-bound, 34.7 clk in
// isolation, profiles/r02_i8_collector_probe.log), every panel byte is read from HBM/L2 half as often per candidate
// (16 row-block pairs instead of 32 row blocks), and the builders' FP64 work per SM and block halves.
//
// Roles per CTA (384 threads, as in the one-CTA kernel): warps 0-3 drain + epilogue, warp 4 TMA producer, warp 5 MMA issuer
// (rank 0 only), warps 8-11 panel builders.  The panel of block j + 1 is built BEFORE the MMAs of block j start, by the drain
// and builder groups together (8 warps): FP64 instructions and kind::i8 MMAs throttle each other (tools/i8_probe2.cu,
// interference runs), so building underneath the MMAs cost both sides more than taking turns does.  Cross-CTA protocol (all cluster-scope release/acquire):
//   full[s]      leader only: the stage tiles of BOTH CTAs (each its own row block of L^-1 + its own panel half, in its own
//                shared memory) have landed -- the copies are 2-D tensor TMA loads with .cta_group::2, whose completion may be
//                signalled on the partner's barrier (plain cp.async.bulk cannot: with the leader's barrier as operand the copy
//                never completes there, and a relay lane in the partner cost 17 % in exposed latency on the two-stage ring:
//                profiles/r02_i8_pair_bringup.log)
//   empty[s]     both: tcgen05.commit of the leader's MMAs, multicast to both CTAs
//   tfull        both: commit multicast when a row-block pair is complete; tempty  leader only, 8 arrivals (4 drain warps x 2 CTAs)
//   xfull[p]     per CTA: the partner's 32 partial column sums for this CTA's candidates have been stored (DSMEM)
// Every wait is bounded (trap, never a hung GPU).  Replaces the same reference code as sweep_kernel
// (optimization/Bayesian7.py:664-682, Bayesian.py:105-112).

#include <cuda.h>     // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)
constexpr int P8_STAGES = BO_P8_STAGES;
constexpr int P8_BH = I8_BN / 2;                       // candidates per CTA (its half of the B operand)
constexpr int P8_B_SLICE = P8_BH * I8_KC;              // 2 KB

template <int S, int DP>
struct P8Smem {
    static constexpr int STAGE_BYTES = S * (I8_A_SLICE + P8_B_SLICE);
    // ring depth: as many stages as 227 KB hold next to ~9 KB of bookkeeping -- three 70 KB stages with 7 slices (the 56 MMAs of a
    // stage last ~1.8 k cycles: a two-stage ring left the issuer waiting for tiles 19 % of the time), two 80 KB stages with 8
    static constexpr int STAGES    = (BO_I8_KC == 64) ? ((3 * STAGE_BYTES + 12 * 1024 <= 232448) ? 3 : 2) : P8_STAGES;
    static constexpr int OFF_BAR   = STAGES * STAGE_BYTES;              // full[] empty[] tfull tempty pfull[2] pempty[2] xfull[2], tmem base
    static constexpr int OFF_COL   = OFF_BAR + 256;                     // colsum[2][4][64] (second set: the SVGP factor's rows)
    static constexpr int OFF_XCH   = OFF_COL + 2 * 4 * I8_BN * 8;       // xch[2][2][32]: partner's partial sums for my candidates
    static constexpr int OFF_MU    = OFF_XCH + 4 * P8_BH * 8;           // mu[2][4][32]: per panel buffer and row quarter
    static constexpr int OFF_PRI   = OFF_MU + 8 * P8_BH * 8;            // pri[2][32] prior variance k(x*, x*), ebc[2][32] operand bound 2^e per candidate
    static constexpr int OFF_TAB   = OFF_PRI + 4 * P8_BH * 8;           // 2^(j/16), j < 16 (exp_neg_fast)
    static constexpr int OFF_TKV   = OFF_TAB + 16 * 8;
    static constexpr int OFF_TKI   = OFF_TKV + BO_MAX_TOPK * 8;
    static constexpr int OFF_ACQ   = OFF_TKI + BO_MAX_TOPK * 8;
    static constexpr int OFF_CMASK = OFF_ACQ + P8_BH * 8;
    static constexpr int OFF_SOB   = OFF_CMASK + 32;
    static constexpr int BYTES     = (OFF_SOB + (BO_MAX_DIM * BO_SOBOL_BITS + BO_MAX_DIM) * 4 + 127) / 128 * 128;
    // The builders' X~ / alpha staging ALIASES the stage ring: a panel is built while the ring is idle (block j + 1 before the
    // MMAs of block j start, after those of block j - 1 have been drained), never while copies are in flight
    static constexpr int OFF_X     = 0;
    static constexpr int X_ROW     = (DP + 2 + 1) * 8;
    static constexpr int X_FREE    = STAGES * STAGE_BYTES;
    static constexpr int XCH       = X_FREE >= 2 * 512 * X_ROW ? 512 : X_FREE >= 2 * 256 * X_ROW ? 256 : 128;     // multiple of 128 (build loop)
    static_assert(X_FREE >= 2 * XCH * X_ROW, "no room for the X~ staging buffers");
    static_assert(BYTES <= 232448, "shared memory");
};

// ---- cluster-scope primitives ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t p8_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t p8_mapa(const void* local, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(rank));
    return ra;
}
__device__ __forceinline__ void p8_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void p8_arrive_cluster(uint64_t* bar) {       // local barrier, cluster-scope release (remote waiters' data)
    asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void p8_st_remote_f64(uint32_t cluster_addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(cluster_addr), "d"(v) : "memory");
}
__device__ __forceinline__ bool p8_try_wait(uint64_t* bar, uint32_t parity) {          // cluster-scope acquire
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
#ifndef BO_I8_WAIT_CYCLES
#define BO_I8_WAIT_CYCLES 100000000000LL
#endif
// CLUSTER = true: cluster-scope acquire (the waiter goes on to read data another CTA stored: the DSMEM exchange; ptxas
// follows every such try_wait with CCTL.IVALL, an L1 invalidate -- 21.8 M of them per 4 ms in the first version, when every
// wait was cluster-scope).  CLUSTER = false: the usual CTA-scope wait for barriers completed by TMA, tcgen05.commit or this
// CTA's own threads (what CUTLASS's 2-SM kernels use for the same barriers).
// NAP > 0: nanoseconds to sleep between polls -- for waits that last most of a block and are not on the critical path (the
// builder group waiting for its panel buffer, the drain warps waiting for the next accumulators): a polling warp still
// issues, and the step is power-bound.
template <bool CLUSTER = false, int NAP = 0>
__device__ __forceinline__ void p8_wait(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    for (uint32_t spin = 0;; ++spin) {
        if (CLUSTER ? p8_try_wait(bar, parity) : mbar_try_wait(bar, parity)) return;
        if (NAP > 0) __nanosleep(NAP);
        if ((spin & 1023u) == 1023u && clock64() - t0 > BO_I8_WAIT_CYCLES) {
            printf("sweep_i8_pair_kernel: mbarrier wait timed out (block %d thread %d barrier +%d parity %u)\n", blockIdx.x, threadIdx.x,
                   (int)(smem_u32(bar) & 0x7f), parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void p8_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int COLL>
__device__ __forceinline__ void p8_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
#define BO_P8_MMA(coll) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8" coll " [%0], %1, %2, %3, p;\n\t}\n" \
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory")
    if (COLL == 1) BO_P8_MMA(".collector::a::fill");
    else if (COLL == 2) BO_P8_MMA(".collector::a::use");
    else if (COLL == 3) BO_P8_MMA(".collector::a::lastuse");
    else BO_P8_MMA("");
#undef BO_P8_MMA
}
// 2-D tiled TMA load of `rows` x 256 B starting at row y of the byte buffer the map describes; lands in THIS CTA's shared
// memory, completes (bytes) on the barrier at cluster address `bar_cluster` -- the leader's
__device__ __forceinline__ void p8_tma_rows(void* smem_dst, const CUtensorMap* tm, uint32_t y, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(0u), "r"(y), "r"(bar_cluster) : "memory");
}
// the same copy with an L2 eviction-priority hint (A/B runs only: flags bits 3-6, see the producer)
__device__ __forceinline__ void p8_tma_rows_hint(void* smem_dst, const CUtensorMap* tm, uint32_t y, uint32_t bar_cluster, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(0u), "r"(y), "r"(bar_cluster), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t p8_policy(bool last) {
    uint64_t pol;
    if (last) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void p8_commit_both(uint64_t* bar) {      // arrives on `bar` of BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <int DP, int KIND, int S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS, 1) sweep_i8_pair_kernel(const SweepArgs a, const SweepI8Args b, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    using SM = P8Smem<S, DP>;
    static_assert(S * I8_BN <= 512, "TMEM has 512 columns");
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full     = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    constexpr int NST = SM::STAGES;
    uint64_t* empty    = full + NST;
    uint64_t* tfull    = empty + NST;
    uint64_t* tempty   = tfull + 1;
    uint64_t* pfull    = tempty + 1;          // [2] panel buffer p holds a finished block half
    uint64_t* pempty   = pfull + 2;           // [2] panel buffer p has been consumed
    uint64_t* xfull    = pempty + 2;          // [2] partner's partial sums have arrived
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(xfull + 2);
    static_assert((2 * NST + 8) * 8 + 4 <= 256, "barrier block");
    double* colsum  = reinterpret_cast<double*>(smem + SM::OFF_COL);
    double* xch     = reinterpret_cast<double*>(smem + SM::OFF_XCH);      // [2][32]
    double* mu_s    = reinterpret_cast<double*>(smem + SM::OFF_MU);       // [2][4][32]
    double* pri_s   = reinterpret_cast<double*>(smem + SM::OFF_PRI);      // [2][32]
    double* ebc_s   = pri_s + 2 * P8_BH;                                  // [2][32]
    double* exp_tab = reinterpret_cast<double*>(smem + SM::OFF_TAB);
    double* tkv     = reinterpret_cast<double*>(smem + SM::OFF_TKV);
    long long* tki  = reinterpret_cast<long long*>(smem + SM::OFF_TKI);
    double* acq_s   = reinterpret_cast<double*>(smem + SM::OFF_ACQ);
    unsigned* cmask = reinterpret_cast<unsigned*>(smem + SM::OFF_CMASK);
    uint32_t* dirs  = reinterpret_cast<uint32_t*>(smem + SM::OFF_SOB);
    uint32_t* shift = dirs + BO_MAX_DIM * BO_SOBOL_BITS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const uint32_t rank = p8_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int nbm = a.np / SW_BM;
    const int nbp = (nbm + 1) / 2;                            // row-block pairs
    // SVGP predictive state (b.sv): the factor is the stack [L^-1; B], B = Ls^T L^-1 -- var = k** + sv_add - ||L^-1 k*||^2 + ||B k*||^2
    // from ONE pass over the same sliced panel (the FP64 kernel's second triangular pass needs u in full first)
    const int nbp_all = b.sv ? 2 * nbp : nbp;
    constexpr int KCH = SW_BM / I8_KC;                        // stages per 128 columns
    constexpr int B_STAGE = S * P8_B_SLICE;
    const size_t tiles_tri = (size_t)nbm * (nbm + 1) / 2 * KCH;               // packed L^-1 tiles; behind them B's (SVGP), then the zero tile
    const size_t tile_zero = tiles_tri + (b.sv ? (size_t)nbm * nbm * KCH : 0);
    const size_t panel_bytes = (size_t)(a.np / I8_KC) * B_STAGE;
    int8_t* panel0 = b.panel8 + (size_t)blockIdx.x * 2 * panel_bytes;      // two half-panel buffers per CTA

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1); mbar_init(tempty, 8);
        for (int s = 0; s < 2; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], 1); mbar_init(&xfull[s], P8_BH); }
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
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
    p8_cluster_sync();                                        // both CTAs' barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    // ---- panel build of block `blk` (the j-th block of this pair) into panel buffer j & 1: K(X, X*) digits of this CTA's 32
    // candidates + their posterior-mean partials.  Run by EIGHT warps -- the drain group (warps 0-3) and the builder group
    // (warps 8-11) together, wb8 = 0..7 -- while the tensor pipe is idle: FP64 instructions crawl next to back-to-back
    // kind::i8 MMAs (one DFMA per ~45 cycles per sub-partition) and slow the MMAs down in turn, so the build of block j + 1 is
    // finished before the MMAs of block j start (see the producer) instead of running underneath them.
    auto build = [&](long long blk, int j, int wb8) {
        const int p = j & 1;
        const int tb8 = wb8 * 32 + lane;
        int8_t* panel = panel0 + (size_t)p * panel_bytes;
        // thread layout (8 warps): warps (wb8 & 1) split the 32 candidates into two groups of 16 (two candidates per thread:
        // 8 g-lanes x gi), warps (wb8 >> 1) split every 128 rows into quarters; the 4 q-lanes of a candidate take 4
        // consecutive rows each
        const int cgrp = wb8 & 1, quarter = wb8 >> 1;
        double xc[2][DP];
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
            long long li = blk * I8_BN + rank * P8_BH + cgrp * 16 + gi * 8 + g;
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
        // linear + Matern kind: k* = s2 (sum_k lin_w[k] x~*_k x~_jk + matern) is not bounded by the output scale; by
        // Cauchy-Schwarz |k*_j| <= s2 (|x*|_w max_j |x_j|_w + 1) =: B, and the candidate's fixed-point scale is the power of
        // two eb above B (its digits, its column of the accumulators and its guard all use that eb)
        constexpr bool LIN = KIND == BO_KERNEL_LINEAR_MATERN52;
        double xw[2][LIN ? DP : 1], dsc[2] = {b.dig_scale, b.dig_scale};
        if (LIN) {
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                double nn = 0.0;
#pragma unroll
                for (int k = 0; k < DP; ++k) { xw[gi][k] = __dmul_rn(a.hyp.lin_w[k], xc[gi][k]); nn = fma(xw[gi][k], xc[gi][k], nn); }
                const double bound = a.hyp.outputscale * (sqrt(nn * __ldg(b.guard_w + 1)) + 1.0);
                const int ec = ((__double2hiint(bound) >> 20) & 0x7ff) - 1023 + 1;              // 2^ec > bound
                dsc[gi] = __hiloint2double((1023 + I8Dig<S>::F - ec) << 20, 0);
                if (quarter == 0 && q == 0) {
                    pri_s[p * P8_BH + cgrp * 16 + gi * 8 + g] = a.hyp.outputscale * (nn + 1.0);
                    ebc_s[p * P8_BH + cgrp * 16 + gi * 8 + g] = __hiloint2double((1023 + ec) << 20, 0);
                }
            }
        }
        double mu0 = 0.0, mu1 = 0.0;
        constexpr int XCH = SM::XCH, XP = DP + 2, XBUF = XCH * XP + XCH;
        double* xstage = reinterpret_cast<double*>(smem + SM::OFF_X);
        const int nrows = a.np;
        const int nchunks = (nrows + XCH - 1) / XCH;
        // X~ rows are staged PERMUTED: row r of a chunk sits at slot (r % 4) * (XCH / 4) + r / 4, so that the rows
        // 4 q + e the four q-lanes of a candidate read in one instruction are adjacent slots (80-byte pitch: four
        // different 16-byte bank groups) -- read in natural order they are 4 rows = 320 B apart, a 2-way bank conflict on
        // every LDS.128 (42 M conflict wavefronts per 19 k candidates in the first version: profiles/r02_ncu_i8_pair_raw.csv)
        auto load_chunk = [&](int c) {
            double* xb = xstage + (c & 1) * XBUF;
            double* ab = xb + XCH * XP;
            const int r0 = c * XCH, rows = min(XCH, nrows - r0);
            for (int e = tb8; e < rows * (DP / 2); e += 256) {
                const int r = e / (DP / 2), k = e % (DP / 2);
                cp_async16(xb + ((r & 3) * (XCH / 4) + (r >> 2)) * XP + 2 * k, a.Xs + (size_t)(r0 + r) * BO_MAX_DIM + 2 * k);
            }
            for (int e = tb8; e < rows / 2; e += 256) cp_async16(ab + 2 * e, a.alpha + r0 + 2 * e);
            cp_async_commit();
        };
        load_chunk(0);
        for (int c = 0; c < nchunks; ++c) {
            if (c + 1 < nchunks) { load_chunk(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const double* xb = xstage + (c & 1) * XBUF;
            const double* ab = xb + XCH * XP;
            const int jend = min(nrows, (c + 1) * XCH);
            // 128 rows per iteration, this warp's quarter of them: the lane evaluates rows j0 + 32 quarter + 16 h + 4 q + e
            // (h < 2, e < 4) for its two candidates -- 16 independent kernel evaluations in flight, one 32-bit word (4
            // consecutive k) of every slice per (candidate, h)
            for (int j0 = c * XCH; j0 < jend; j0 += 128) {
                double kv[2][8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int j = j0 + 32 * quarter + 16 * (r >> 2) + 4 * q + (r & 3);
                    const int jr = j - c * XCH;
                    const double2* row = reinterpret_cast<const double2*>(xb + ((jr & 3) * (XCH / 4) + (jr >> 2)) * XP);
                    double x[DP];
#pragma unroll
                    for (int k = 0; k < DP / 2; ++k) { const double2 t = row[k]; x[2 * k] = t.x; x[2 * k + 1] = t.y; }
                    const double al = ab[jr];
#pragma unroll
                    for (int gi = 0; gi < 2; ++gi) {
                        double sq = 0.0;
#pragma unroll
                        for (int k = 0; k < DP; ++k) { const double df = __dsub_rn(xc[gi][k], x[k]); sq = fma(df, df, sq); }
                        double v = kernel_value_fast_t<LIN ? BO_KERNEL_MATERN52 : KIND>(sq, a.hyp.outputscale, exp_tab);
                        if (LIN) {
                            double lin = 0.0;
#pragma unroll
                            for (int k = 0; k < DP; ++k) lin = fma(xw[gi][k], x[k], lin);
                            v = fma(a.hyp.outputscale, lin, v);
                        }
                        kv[gi][r] = (j < a.n) ? v : 0.0;
                    }
                    mu0 = fma(kv[0][r], al, mu0); mu1 = fma(kv[1][r], al, mu1);
                }
#pragma unroll
                for (int gi = 0; gi < 2; ++gi)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        // 16-row chunk (2 quarter + hh) of the 128 rows: its stage tile and K chunk within the tile
                        const int jj = j0 + 32 * quarter + 16 * hh;
                        int8_t* st_tile = panel + (size_t)(jj / I8_KC) * B_STAGE;
                        const int ch = (jj % I8_KC) / 16;
                        uint32_t w[S];
                        long long v4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) v4[e] = __double2ll_rn(kv[gi][hh * 4 + e] * dsc[gi]);
                        i8_digit_words<S>(v4, w);
                        const size_t off = ((size_t)(cgrp * 2 + gi) * (I8_KC / 16) + ch) * 128 + g * 16 + q * 4;
#pragma unroll
                        for (int s = 0; s < S; ++s) *reinterpret_cast<uint32_t*>(st_tile + (size_t)s * P8_B_SLICE + off) = w[s];
                    }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");     // buffer free before chunk c + 2 overwrites it
        }
        mu0 += __shfl_xor_sync(0xffffffffu, mu0, 1); mu0 += __shfl_xor_sync(0xffffffffu, mu0, 2);
        mu1 += __shfl_xor_sync(0xffffffffu, mu1, 1); mu1 += __shfl_xor_sync(0xffffffffu, mu1, 2);
        if (q == 0) {             // per row quarter: the epilogue adds the four in a fixed order
            mu_s[(p * 4 + quarter) * P8_BH + cgrp * 16 + g] = mu0;
            mu_s[(p * 4 + quarter) * P8_BH + cgrp * 16 + 8 + g] = mu1;
        }
        __threadfence();
        fence_proxy_async();      // generic-proxy panel writes -> visible to the async-proxy (TMA) reads
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tb8 == 0) mbar_arrive(&pfull[p]);
    };

    if (warp >= 8) {
        // ================= builder group: with the drain group, builds block j once its panel buffer is free =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;" ::: "memory");
        int j = 0;
        for (long long blk = pair; blk < a.nblocks; blk += npairs, ++j) {
            p8_wait<false, 500>(&pempty[j & 1], ((j >> 1) & 1) ^ 1);      // the block that used this buffer two turns ago is done (epilogue)
            build(blk, j, 4 + (warp - 8));
        }
    } else if (warp >= 4) {
        // ================= warpgroup 1: TMA producer (warp 4), MMA issuer / relay (warp 5) =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;" ::: "memory");
        if (warp < 6) {
        const bool prof = (a.flags & 4) != 0 && rank == 0;
        const bool profp = (a.flags & 4) != 0;            // producers of both ranks: time spent waiting for a free slot
        long long t_pe = 0, t_pp = 0;
        long long t_tot = 0, t_w0 = 0, t_w1 = 0;
        unsigned long long ns0 = 0, ns1 = 0;
        if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
        const long long c0 = prof ? clock64() : 0;
        const long long c0p = profp ? clock64() : 0;
        int stage = 0; uint32_t phase = 0;
        uint32_t rb = 0;                          // running row-block-pair counter (accumulator full/empty phases)
        int it = 0;
        // L2 eviction-priority experiments (BO_B200_SWEEP_FLAGS bits: 8 / 16 = L^-1 tiles evict_last / evict_first, 32 / 64 = panel tiles)
        const bool hintA = (a.flags & 24) != 0, hintB = (a.flags & 96) != 0;
        const uint64_t polA = hintA ? p8_policy((a.flags & 8) != 0) : 0, polB = hintB ? p8_policy((a.flags & 32) != 0) : 0;
        for (long long blk = pair; blk < a.nblocks; blk += npairs, ++it) {
            const int p = it & 1;
            const int8_t* panel = panel0 + (size_t)p * panel_bytes;
            const long long tB0 = prof ? clock64() : 0;
            if (warp == 4) {
                if (lane == 0) {
                    const long long wq = profp ? clock64() : 0;
                    p8_wait(&pfull[p], (it >> 1) & 1);        // this CTA's builders have finished their half of the block
                    // FP64 instructions crawl while the INT8 tensor pipe is busy (a DFMA chain issues once per ~45 cycles next
                    // to back-to-back kind::i8 MMAs, and each one costs the tensor pipe ~1 cycle: profiles/r02_i8_collector_probe.log,
                    // interference runs), so the panel of the NEXT block is built before this block's MMAs start rather than
                    // underneath them: the builders run at full FP64 rate and the MMAs undisturbed
                    // (not optional: the builders stage X~ through the idle stage ring)
                    if (blk + npairs < a.nblocks) p8_wait(&pfull[p ^ 1], ((it + 1) >> 1) & 1);
                    if (profp) t_pp += clock64() - wq;
                    for (int ibp = 0; ibp < nbp_all; ++ibp) {
                        // SVGP: after the row blocks of L^-1 (lower block triangle) come those of B = Ls^T L^-1 (dense: every K step)
                        const bool dense = ibp >= nbp;
                        const int ib = 2 * (dense ? ibp - nbp : ibp) + (int)rank;   // this CTA's row block of the pair
                        const int nkc = dense ? nbm * KCH : (2 * ibp + 2) * KCH;    // the pair walks the K extent of its odd row block
                        const int mykc = ib < nbm ? (dense ? nbm * KCH : (ib + 1) * KCH) : 0;
                        const size_t tile0 = dense ? tiles_tri + (size_t)ib * nbm * KCH : (size_t)ib * (ib + 1) / 2 * KCH;
                        for (int kc = 0; kc < nkc; ++kc) {
                            const long long we = profp ? clock64() : 0;
                            p8_wait(&empty[stage], phase ^ 1);
                            if (profp) t_pe += clock64() - we;
                            unsigned char* sb = smem + stage * SM::STAGE_BYTES;
                            if (rank == 0) mbar_expect_tx(&full[stage], 2u * SM::STAGE_BYTES);     // both CTAs' tiles complete here
                            // beyond this row block's own K extent (the even block's last 128 columns, a row block past the
                            // end): structurally zero -> the all-zero tile parked behind the packed factor
                            const size_t tile = kc < mykc ? tile0 + kc : tile_zero;
                            const uint32_t fb = p8_mapa(&full[stage], 0);
                            if (hintA) p8_tma_rows_hint(sb, &tmA, (uint32_t)(tile * (S * I8_A_SLICE / 256)), fb, polA);
                            else       p8_tma_rows(sb, &tmA, (uint32_t)(tile * (S * I8_A_SLICE / 256)), fb);
                            const uint32_t yb = (uint32_t)((panel - b.panel8 + (size_t)kc * B_STAGE) / 256);
                            if (hintB) p8_tma_rows_hint(sb + S * I8_A_SLICE, &tmB, yb, fb, polB);
                            else       p8_tma_rows(sb + S * I8_A_SLICE, &tmB, yb, fb);
                            if (++stage == NST) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            } else if (rank != 0) {
                // the partner's MMA warp has nothing to do: the leader issues for both CTAs
            } else {
                // leader: one elected lane issues the M = 256 MMAs of both CTAs (fully unrolled, warp-uniform control flow)
                const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_BN >> 3) << 17) | ((uint32_t)((2 * SW_BM) >> 4) << 24);
                const bool leader = i8_elect();
                for (int ibp = 0; ibp < nbp_all; ++ibp, ++rb) {
                    const long long w1 = prof ? clock64() : 0;
                    p8_wait<true>(tempty, (rb & 1) ^ 1);       // both CTAs have drained the previous row-block pair
                    if (prof) t_w1 += clock64() - w1;
                    tc_fence_after();
                    const int nkc = ibp >= nbp ? nbm * KCH : (2 * ibp + 2) * KCH;
                    for (int kc = 0; kc < nkc; ++kc) {
                        const long long w0 = prof ? clock64() : 0;
                        p8_wait(&full[stage], phase);
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
                                        const uint64_t db = db0 + (uint64_t)((t * P8_B_SLICE + kk * 256) >> 4);
                                        const uint32_t accf = (kk > 0 || s > 0) ? 1u : acc0;
                                        if (s == S - 1)          p8_mma<0>(td, da, db, idesc, accf);
                                        else if (t == 0)         p8_mma<1>(td, da, db, idesc, accf);
                                        else if (t == S - 1 - s) p8_mma<3>(td, da, db, idesc, accf);
                                        else                     p8_mma<2>(td, da, db, idesc, accf);
                                    }
                            p8_commit_both(&empty[stage]);               // both CTAs' slots are free once these MMAs have read them
                            if (kc == nkc - 1) p8_commit_both(tfull);    // the accumulators of the row-block pair are complete
                        }
                        __syncwarp();
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                }
            }
            if (prof && warp == 5) t_tot += clock64() - tB0;
        }
        if (profp && blockIdx.x < 2 && tid == 128)
            printf("sweep_i8_pair CTA %d TMA producer: waits: slot-empty %lld clk, panel-ready %lld clk (of %lld)\n", blockIdx.x, t_pe, t_pp, clock64() - c0p);
        if (prof && blockIdx.x == 0 && tid == 160) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
            const long long c1 = clock64();
            printf("sweep_i8_pair CTA 0 MMA issuer: %d blocks, %lld clk in the contraction; waits: stage-full (both CTAs) %lld, "
                   "accumulators-drained %lld; SM clock over the kernel %.0f MHz\n", it, t_tot, t_w0, t_w1, (double)(c1 - c0) / (double)(ns1 - ns0) * 1e3);
        }
        }
    } else {
        // ================= warpgroup 0: accumulator drain (own row block) + epilogue (own 32 candidates) =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
        uint32_t rb = 0;
        int it = 0;
        const uint32_t tempty_leader = p8_mapa(tempty, 0);
        // the first two panels, together with the builder group
        build(pair, 0, warp);
        if ((long long)pair + npairs < a.nblocks) build((long long)pair + npairs, 1, warp);
        for (long long blk = pair; blk < a.nblocks; blk += npairs, ++it) {
            const int p = it & 1;
            for (int ph = 0; ph < (b.sv ? 2 : 1); ++ph) {          // ph = 1: the rows of B (SVGP), summed apart from those of L^-1
                double acc[I8_BN];
#pragma unroll
                for (int c = 0; c < I8_BN; ++c) acc[c] = 0.0;
                for (int ibp = 0; ibp < nbp; ++ibp, ++rb) {
                    const int ib = 2 * ibp + (int)rank;
                    const double rs = ib < nbm ? b.rowscale[ph * a.np + ib * SW_BM + tid] * b.eb_scale : 0.0;
                    p8_wait<false, 100>(tfull, rb & 1);
                    tc_fence_after();
                    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
                    for (int c0 = 0; c0 < I8_BN; c0 += 8) {
                        int v[S][8];
#pragma unroll
                        for (int gq = 0; gq < S; ++gq) tmem_ld8(trow + gq * I8_BN + c0, v[gq]);
                        tmem_ld_wait();
#pragma unroll
                        for (int gq = 0; gq < S; ++gq)
                            asm volatile("" : "+r"(v[gq][0]), "+r"(v[gq][1]), "+r"(v[gq][2]), "+r"(v[gq][3]), "+r"(v[gq][4]), "+r"(v[gq][5]), "+r"(v[gq][6]), "+r"(v[gq][7]));
                        if (c0 + 8 == I8_BN) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) p8_arrive_remote(tempty_leader);
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
                {
                    int base = 0;
#pragma unroll
                    for (int o = 16, cnt = I8_BN / 2; o >= 1; o >>= 1, cnt >>= 1) base += (lane & o) ? cnt : 0;
                    colsum[(ph * 4 + warp) * I8_BN + base] = acc[0];
                    colsum[(ph * 4 + warp) * I8_BN + base + 1] = acc[1];
                }
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");

            // partial sums over this CTA's row blocks: the partner's candidates go to the partner, mine stay
            double mine = 0.0, mine2 = 0.0;
            if (tid < P8_BH) {
                const int cp = (1 - (int)rank) * P8_BH + tid, cm = (int)rank * P8_BH + tid;
                const double theirs = (colsum[cp] + colsum[I8_BN + cp]) + (colsum[2 * I8_BN + cp] + colsum[3 * I8_BN + cp]);
                mine = (colsum[cm] + colsum[I8_BN + cm]) + (colsum[2 * I8_BN + cm] + colsum[3 * I8_BN + cm]);
                p8_st_remote_f64(p8_mapa(&xch[p * P8_BH + tid], 1 - rank), theirs);
                if (b.sv) {
                    const double* c2 = colsum + 4 * I8_BN;
                    const double theirs2 = (c2[cp] + c2[I8_BN + cp]) + (c2[2 * I8_BN + cp] + c2[3 * I8_BN + cp]);
                    mine2 = (c2[cm] + c2[I8_BN + cm]) + (c2[2 * I8_BN + cm] + c2[3 * I8_BN + cm]);
                    p8_st_remote_f64(p8_mapa(&xch[(2 + p) * P8_BH + tid], 1 - rank), theirs2);
                }
                p8_arrive_remote(p8_mapa(&xfull[p], 1 - rank));           // release.cluster: orders this thread's store before it
                p8_wait<true>(&xfull[p], (it >> 1) & 1);
                p8_wait(&pfull[p], (it >> 1) & 1);                        // acquire: mu_s[p] of this CTA's builders
            }
            // ================= epilogue: variance, acquisition, CTA-local top-k (32 candidates) =====================
            if (tid < P8_BH) {
                const long long li = blk * I8_BN + rank * P8_BH + tid;
                const double other = xch[p * P8_BH + tid];
                constexpr bool LIN = KIND == BO_KERNEL_LINEAR_MATERN52;
                const double ebc = LIN ? ebc_s[p * P8_BH + tid] : 1.0;           // (stationary kinds: eb is folded into ss_scale / guard_scale)
                const double prior = LIN ? pri_s[p * P8_BH + tid] : a.hyp.outputscale;
                // (even row blocks) + (odd row blocks): one order for both CTAs; then the operand bound squared (a power of two)
                const double ssu = (rank == 0 ? mine + other : other + mine) * (LIN ? ebc * ebc : b.ss_scale);
                double ss = ssu, gerr = sqrt(__ldg(b.guard_w) * ssu);
                if (b.sv) {                 // ||B k*||^2 comes back; its slicing error adds to the bound (W of B's rows: guard_w[2])
                    const double other2 = xch[(2 + p) * P8_BH + tid];
                    const double ssw = (rank == 0 ? mine2 + other2 : other2 + mine2) * (LIN ? ebc * ebc : b.ss_scale);
                    ss = ssu - ssw - b.sv_add;
                    gerr += sqrt(__ldg(b.guard_w + 2) * ssw);
                }
                const bool flagged = b.flag_count != nullptr && li < a.N && !(prior - ss >= b.guard_scale * ebc * gerr);
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
                const double var = fmax(prior - ss, a.min_var);
                const double mean = a.hyp.mean + ((mu_s[(p * 4) * P8_BH + tid] + mu_s[(p * 4 + 1) * P8_BH + tid]) +
                                                  (mu_s[(p * 4 + 2) * P8_BH + tid] + mu_s[(p * 4 + 3) * P8_BH + tid]));
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
                if (lane == 0) cmask[0] = m;
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (tid == 0) {
                mbar_arrive(&pempty[p]);           // panel half p and mu_s[p] are free for the builders (block it + 2)
                if (a.topk > 0) {
                    const int K = a.topk;
                    unsigned m = cmask[0];
                    while (m) {
                        const int c = __ffs(m) - 1;
                        m &= m - 1;
                        const double v = acq_s[c];
                        const long long gi = a.first_index + blk * I8_BN + rank * P8_BH + c;
                        if (!tk_better(v, gi, tkv[K - 1], tki[K - 1])) continue;
                        int pp = K - 1;
                        while (pp > 0 && tk_better(v, gi, tkv[pp - 1], tki[pp - 1])) { tkv[pp] = tkv[pp - 1]; tki[pp] = tki[pp - 1]; --pp; }
                        tkv[pp] = v; tki[pp] = gi;
                    }
                }
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            // buffer p is free again (pempty above): help build the block after next into it -- its MMAs wait for that
            if (blk + 2LL * npairs < a.nblocks) build(blk + 2LL * npairs, it + 2, warp);
        }
        if (tid < BO_MAX_TOPK && a.part_val) {
            a.part_val[(size_t)blockIdx.x * BO_MAX_TOPK + tid] = tkv[tid];
            a.part_idx[(size_t)blockIdx.x * BO_MAX_TOPK + tid] = tki[tid];
        }
    }
    tc_fence_before();
    __syncthreads();
    p8_cluster_sync();                 // neither CTA frees tensor memory (or exits with live barriers) while the partner still uses it
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// Tensor maps over the packed factor (+ zero tile) and the panel buffer, both viewed as rows of 256 bytes; one box = one
// stage tile (S slices).  The encoder is a host-side driver entry point.
static int p8_encode_rows(bo_handle* h, CUtensorMap* tm, const void* base, size_t bytes, uint32_t box_rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        BO_CUDA(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn || qr != cudaDriverEntryPointSuccess) return fail(h, BO_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t dims[2] = {256, (cuuint64_t)(bytes / 256)};
    const cuuint64_t strides[1] = {256};
    const cuuint32_t box[2] = {256, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, BO_E_CUDA, "cuTensorMapEncodeTiled failed for the sliced sweep's operand buffers");
    return 0;
}

template <int DP, int KIND, int S>
static int launch_sweep_i8_pair_k(bo_handle* h, const SweepArgs& a, const SweepI8Args& b, int grid, cudaStream_t st) {
    alignas(64) CUtensorMap tmA, tmB;
    int rc;
    if ((rc = p8_encode_rows(h, &tmA, h->Lp8, h->Lp8_bytes, S * I8_A_SLICE / 256))) return rc;
    if ((rc = p8_encode_rows(h, &tmB, h->panel8, h->panel8_bytes, S * P8_B_SLICE / 256))) return rc;
    BO_CUDA(h, cudaFuncSetAttribute(sweep_i8_pair_kernel<DP, KIND, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, P8Smem<S, DP>::BYTES));
    sweep_i8_pair_kernel<DP, KIND, S><<<grid, I8_THREADS, P8Smem<S, DP>::BYTES, st>>>(a, b, tmA, tmB);     // cluster dims are compiled in
    BO_LAUNCH_CHECK(h);
    return 0;
}
template <int DP>
static int launch_sweep_i8_pair(bo_handle* h, const SweepArgs& a, const SweepI8Args& b, int S, int grid, cudaStream_t st) {
    if (a.hyp.kind == BO_KERNEL_MATERN52)
        return S == 8 ? launch_sweep_i8_pair_k<DP, BO_KERNEL_MATERN52, 8>(h, a, b, grid, st) : launch_sweep_i8_pair_k<DP, BO_KERNEL_MATERN52, 7>(h, a, b, grid, st);
    if (a.hyp.kind == BO_KERNEL_LINEAR_MATERN52)
        return S == 8 ? launch_sweep_i8_pair_k<DP, BO_KERNEL_LINEAR_MATERN52, 8>(h, a, b, grid, st)
                      : launch_sweep_i8_pair_k<DP, BO_KERNEL_LINEAR_MATERN52, 7>(h, a, b, grid, st);
    return S == 8 ? launch_sweep_i8_pair_k<DP, BO_KERNEL_RBF, 8>(h, a, b, grid, st) : launch_sweep_i8_pair_k<DP, BO_KERNEL_RBF, 7>(h, a, b, grid, st);
}
