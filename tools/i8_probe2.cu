// Round-2 INT8 tcgen05 probe: what lifts the sliced sweep's 128 x 64 x 32 MMA off its 50-cycle floor (ideal 32)?
//   (a) collector reuse of the A operand (.collector::a::fill / ::use / ::lastuse -> SASS A_KEEP / A_REUSE): in the sweep's
//       issue order slice s of L^-1 multiplies S - s panel slices back to back, so A need only be read S times per 36 MMAs;
//   (b) N = 128 tiles (accumulator groups folded: rate only);
//   (c) CTA pairs (cta_group::2, M = 256): each SM reads half of B.
// Every variant is first checked bit-exactly against an integer reference (when all S groups fit TMEM), then timed in
// SM cycles with the fully unrolled, warp-uniform issue loop the kernel uses.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/i8_probe2 tools/i8_probe2.cu && timeout 300 tools/i8_probe2
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__host__ __device__ inline int digit(uint32_t slice, uint32_t which, uint32_t r, uint32_t k) {
    uint32_t h = (slice * 0x9E3779B1u) ^ (which * 0x85EBCA77u) ^ (r * 0xC2B2AE3Du) ^ (k * 0x27D4EB2Fu);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return (int)(h % 129u) - 64;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// COLL: 0 plain, 1 fill, 2 use, 3 lastuse
template <int CG, int COLL>
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
#define MMA_ASM(cgs, coll) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::" cgs ".kind::i8" coll " [%0], %1, %2, %3, p;\n\t}\n" \
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory")
    if (CG == 1) {
        if (COLL == 0) MMA_ASM("1", ""); else if (COLL == 1) MMA_ASM("1", ".collector::a::fill");
        else if (COLL == 2) MMA_ASM("1", ".collector::a::use"); else MMA_ASM("1", ".collector::a::lastuse");
    } else {
        if (COLL == 0) MMA_ASM("2", ""); else if (COLL == 1) MMA_ASM("2", ".collector::a::fill");
        else if (COLL == 2) MMA_ASM("2", ".collector::a::use"); else MMA_ASM("2", ".collector::a::lastuse");
    }
#undef MMA_ASM
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// CG: CTAs per MMA (1 or 2; with 2 the launch is a cluster of 2 and the MMA is M = 256, each CTA holding 128 rows of A and
// N / 2 rows of B); N: MMA N; S slices; G accumulator groups kept (G < S folds groups: rate only); COLL: A-collector reuse;
// ORDER: 0 = for s, for t (A slice reused back to back), 1 = for t, for s (B slice reused back to back: control)
// side work for the interference runs: warps 4..7 (one per SM sub-partition) spin on independent chains of one instruction
// class until the issuing warp raises `stop`; EXTRA: 1 = DFMA, 2 = FFMA, 3 = IMAD, 4 = DFMA at 1/4 duty (3 of 4 slots idle)
__device__ double g_sink;
template <int CG, int N, int S, int G, int COLL, int ORDER, int KC, int EXTRA = 0>
__global__ void __launch_bounds__(EXTRA ? 256 : 128, 1)
probe_kernel(int iters, int32_t* __restrict__ out, int* __restrict__ err, long long* __restrict__ cycles) {
    __shared__ volatile int stop;
    if (threadIdx.x == 0) stop = 0;
    constexpr int M = 128;                                  // rows of A per CTA
    constexpr int NB = N / CG;                              // rows of B per CTA
    constexpr int A_TILE = M * KC, B_TILE = NB * KC;
    static_assert(G * N <= 512, "TMEM has 512 columns");
    constexpr uint32_t COLS = (G * N <= 32) ? 32 : (G * N <= 64) ? 64 : (G * N <= 128) ? 128 : (G * N <= 256) ? 256 : 512;
    extern __shared__ __align__(1024) uint8_t smem[];
    int8_t* sA = reinterpret_cast<int8_t*>(smem);
    int8_t* sB = sA + S * A_TILE;
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));

    for (int i = tid; i < S * A_TILE; i += 128) {
        int s = i / A_TILE, o = i % A_TILE, cm = o / 128, w = o % 128;
        sA[i] = (int8_t)digit(s, 0, rank * M + (cm / (KC / 16)) * 8 + w / 16, (cm % (KC / 16)) * 16 + w % 16);
    }
    for (int i = tid; i < S * B_TILE; i += 128) {
        int s = i / B_TILE, o = i % B_TILE, cm = o / 128, w = o % 128;
        sB[i] = (int8_t)digit(s, 1, rank * NB + (cm / (KC / 16)) * 8 + w / 16, (cm % (KC / 16)) * 16 + w % 16);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cg::this_cluster().sync(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((M * CG) >> 4) << 24);
    const uint64_t da0 = make_desc(smem_u32(sA), 128, (KC / 16) * 128), db0 = make_desc(smem_u32(sB), 128, (KC / 16) * 128);
    bool ok = true;
    if (EXTRA && warp >= 4) {
        double d0 = 1.0 + tid, d1 = 2.0, d2 = 3.0, d3 = 4.0, d4 = 5.0, d5 = 6.0, d6 = 7.0, d7 = 8.0;
        float f0 = 1.f + tid, f1 = 2.f, f2 = 3.f, f3 = 4.f, f4 = 5.f, f5 = 6.f, f6 = 7.f, f7 = 8.f;
        int i0 = tid, i1 = 2, i2 = 3, i3 = 4, i4 = 5, i5 = 6, i6 = 7, i7 = 8;
        long long n = 0;
        while (!stop) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                if (EXTRA == 1 || EXTRA == 4) { d0 = fma(d0, 1.0000001, 1e-9); d1 = fma(d1, 1.0000001, 1e-9); d2 = fma(d2, 1.0000001, 1e-9); d3 = fma(d3, 1.0000001, 1e-9);
                                  d4 = fma(d4, 1.0000001, 1e-9); d5 = fma(d5, 1.0000001, 1e-9); d6 = fma(d6, 1.0000001, 1e-9); d7 = fma(d7, 1.0000001, 1e-9); }
                if (EXTRA == 2) { f0 = fmaf(f0, 1.0001f, 1e-5f); f1 = fmaf(f1, 1.0001f, 1e-5f); f2 = fmaf(f2, 1.0001f, 1e-5f); f3 = fmaf(f3, 1.0001f, 1e-5f);
                                  f4 = fmaf(f4, 1.0001f, 1e-5f); f5 = fmaf(f5, 1.0001f, 1e-5f); f6 = fmaf(f6, 1.0001f, 1e-5f); f7 = fmaf(f7, 1.0001f, 1e-5f); }
                if (EXTRA == 3) { i0 = i0 * 3 + 1; i1 = i1 * 3 + 1; i2 = i2 * 3 + 1; i3 = i3 * 3 + 1; i4 = i4 * 3 + 1; i5 = i5 * 3 + 1; i6 = i6 * 3 + 1; i7 = i7 * 3 + 1; }
            }
            if (EXTRA == 4) __nanosleep(0), n += 0;
            if (EXTRA == 4) { for (int w = 0; w < 6; ++w) asm volatile("nanosleep.u32 20;"); }
            n += 128;
        }
        if (d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7 == 12345.678) g_sink = d0;
        if ((tid & 31) == 0 && blockIdx.x == 0 && cycles) cycles[1 + (warp - 4)] = n;       // side instructions issued per lane
    }
    if (warp == 1) {
        uint32_t leader;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(leader));
        const long long c0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t acc0 = it > 0 ? 1u : 0u;
            if (leader && rank == 0) {
#pragma unroll
                for (int kk = 0; kk < KC / 32; ++kk) {
                    if (ORDER == 0) {
#pragma unroll
                        for (int s = 0; s < S; ++s)
#pragma unroll
                            for (int t = 0; t + s < S; ++t) {
                                const int g = (s + t) % G;
                                const bool first = (s + t < G) && s == 0 && kk == 0;     // first touch of accumulator g in this iteration
                                const uint32_t accf = first ? acc0 : 1u;
                                const uint64_t da = da0 + (uint64_t)((s * A_TILE + kk * 256) >> 4), db = db0 + (uint64_t)((t * B_TILE + kk * 256) >> 4);
                                const int run = S - s;
                                if (COLL == 0 || run == 1) mma_i8<CG, 0>(tmem_base + g * N, da, db, idesc, accf);
                                else if (t == 0)           mma_i8<CG, 1>(tmem_base + g * N, da, db, idesc, accf);
                                else if (t == run - 1)     mma_i8<CG, 3>(tmem_base + g * N, da, db, idesc, accf);
                                else                       mma_i8<CG, 2>(tmem_base + g * N, da, db, idesc, accf);
                            }
                    } else {
#pragma unroll
                        for (int t = 0; t < S; ++t)
#pragma unroll
                            for (int s = 0; t + s < S; ++s) {
                                const int g = (s + t) % G;
                                const bool first = (s + t < G) && t == 0 && kk == 0;
                                const uint32_t accf = first ? acc0 : 1u;
                                const uint64_t da = da0 + (uint64_t)((s * A_TILE + kk * 256) >> 4), db = db0 + (uint64_t)((t * B_TILE + kk * 256) >> 4);
                                mma_i8<CG, 0>(tmem_base + g * N, da, db, idesc, accf);
                            }
                    }
                }
                commit<CG>(&bars[it & 1]);
            }
            __syncwarp();
            if (it > 0) ok = ok && mbar_wait_bounded(&bars[(it - 1) & 1], ((it - 1) >> 1) & 1);
            if (!ok) break;
        }
        if (ok) ok = mbar_wait_bounded(&bars[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
        const long long c1 = clock64();
        if (!ok && leader) atomicAdd(err, 1);
        if (leader && blockIdx.x == 0 && cycles) *cycles = c1 - c0;
        stop = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (out != nullptr && blockIdx.x < CG) {
        for (int c0 = 0; c0 < G * N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) out[(size_t)(rank * M + warp * 32 + (tid & 31)) * (G * N) + c0 + j] = (int32_t)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cg::this_cluster().sync(); else __syncthreads();
    if (warp == 0) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(COLS) : "memory");
        else         asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(COLS) : "memory");
    }
}

template <int CG, int N, int S, int G, int COLL, int ORDER, int KC, int EXTRA = 0>
static void run(int sms, const char* what) {
    constexpr int M = 128;
    auto kern = probe_kernel<CG, N, S, G, COLL, ORDER, KC, EXTRA>;
    size_t smem = (size_t)S * (M + N / CG) * KC;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int32_t* d_out; int* d_err; long long* d_cyc;
    const size_t nout = (size_t)M * CG * G * N;
    CK(cudaMalloc(&d_out, sizeof(int32_t) * nout)); CK(cudaMemset(d_out, 0xff, sizeof(int32_t) * nout));
    CK(cudaMalloc(&d_err, sizeof(int))); CK(cudaMemset(d_err, 0, sizeof(int)));
    CK(cudaMalloc(&d_cyc, 8 * sizeof(long long))); CK(cudaMemset(d_cyc, 0, 8 * sizeof(long long)));
    auto launch = [&](int grid, int iters, int32_t* o) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(EXTRA ? 256 : 128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, kern, iters, o, d_err, d_cyc));
    };
    launch(CG, 1, d_out);
    CK(cudaDeviceSynchronize());
    long bad = -1;
    if (G == S) {
        std::vector<int32_t> out(nout);
        CK(cudaMemcpy(out.data(), d_out, nout * sizeof(int32_t), cudaMemcpyDeviceToHost));
        bad = 0;
        for (int g = 0; g < G; ++g)
            for (int r = 0; r < M * CG; ++r)
                for (int c = 0; c < N; ++c) {
                    long ref = 0;
                    for (int s = 0; s <= g; ++s)
                        for (int k = 0; k < KC; ++k) ref += (long)digit(s, 0, r, k) * digit(g - s, 1, c, k);
                    if ((long)out[(size_t)r * (G * N) + g * N + c] != ref) ++bad;
                }
    }
    const int iters = 2000;
    const int grid = sms / CG * CG;
    launch(grid, iters, nullptr);
    CK(cudaDeviceSynchronize());
    long long cyc = 0, side[8] = {0}; int e = 0;
    CK(cudaMemcpy(side, d_cyc, 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    cyc = side[0];
    CK(cudaMemcpy(&e, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    const int mmas = S * (S + 1) / 2 * (KC / 32);
    const double clk = (double)cyc / iters / mmas;
    const double ideal = (double)M * N * 32 / 8192.0;       // 8192 int8 MAC / clk / SM
    printf("%-58s cg%d M=%3d N=%3d S=%d G=%d KC=%d: exact %s; %6.1f clk per MMA (math floor %.0f -> %.2f of the pipe)%s\n", what, CG, M * CG, N, S, G, KC,
           bad < 0 ? "n/a (folded)" : bad == 0 ? "yes" : "NO", clk, ideal, ideal / clk, e ? "  [TIMEOUT]" : "");
    if (bad > 0) printf("    %ld accumulators differ\n", bad);
    if (EXTRA) printf("    side work (warp 4): %.3f instructions per clk per sub-partition; MMA time stolen per side warp-instruction (4 warps): see clk per MMA above\n",
                      (double)side[1] / (double)cyc);
    cudaFree(d_out); cudaFree(d_err); cudaFree(d_cyc);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs, cc %d.%d\n", p.name, p.multiProcessorCount, p.major, p.minor);
    const int sms = p.multiProcessorCount;
    run<1, 64, 8, 8, 0, 0, 64>(sms, "baseline: SS, slice-major (what round 1 ships)");
    run<1, 64, 8, 8, 0, 1, 64>(sms, "control: SS, panel-slice-major order");
    run<1, 64, 8, 8, 1, 0, 64>(sms, "A collector reuse (fill/use/lastuse)");
    run<1, 64, 7, 7, 1, 0, 64>(sms, "A collector reuse, 7 slices");
    run<1, 128, 8, 4, 0, 0, 32>(sms, "N=128 (groups folded), plain");
    run<1, 128, 8, 4, 1, 0, 32>(sms, "N=128 (groups folded), A collector reuse");
    run<1, 256, 8, 2, 0, 0, 32>(sms, "N=256 (groups folded), plain");
    run<2, 64, 8, 8, 0, 0, 64>(sms, "CTA pair, plain");
    run<2, 64, 8, 8, 1, 0, 64>(sms, "CTA pair, A collector reuse");
    run<2, 128, 8, 4, 0, 0, 32>(sms, "CTA pair N=128 (groups folded), plain");
    run<2, 128, 8, 4, 1, 0, 32>(sms, "CTA pair N=128 (groups folded), A collector reuse");
    run<2, 128, 4, 4, 1, 0, 64>(sms, "CTA pair N=128 S=4 exactness, A collector reuse");
    // interference: what does concurrent scalar work on the other warps cost the tensor pipe?
    run<1, 64, 8, 8, 1, 0, 64, 1>(sms, "A reuse + 4 warps of DFMA chains");
    run<1, 64, 8, 8, 1, 0, 64, 4>(sms, "A reuse + 4 warps of DFMA at low duty");
    run<1, 64, 8, 8, 1, 0, 64, 2>(sms, "A reuse + 4 warps of FFMA chains");
    run<1, 64, 8, 8, 1, 0, 64, 3>(sms, "A reuse + 4 warps of IMAD chains");
    run<2, 64, 8, 8, 1, 0, 64, 1>(sms, "CTA pair, A reuse + 4 warps of DFMA chains");
    return 0;
}
