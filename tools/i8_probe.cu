// INT8 tensor-path probe for DESIGN §7 (Ozaki-sliced FP64 contraction): is tcgen05.mma kind::i8 usable from
// hand-written code on sm_100a, is its INT32 accumulation exact, and what rate does it sustain at the tile shapes
// the TMEM budget allows when S slices of A (128 x KC) and S slices of B (N x KC) sit in shared memory and every
// slice pair s+t < S is multiplied (S(S+1)/2 products per K step, products with equal s+t share one accumulator)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/i8_probe tools/i8_probe.cu
//   timeout 120 tools/i8_probe
//
// Shared-memory operand layout: K-major, no swizzle.  A slice tile of R rows x KC bytes is a grid of 8-row x 16-byte
// "core matrices" (128 contiguous bytes each): element (r,k) at ((r/8)*(KC/16) + k/16)*128 + (r%8)*16 + k%16.
// Matrix descriptor: LBO = 128 B (next core matrix along K), SBO = (KC/16)*128 B (next 8-row group).
// Accumulator D (128 x N, s32) lives in TMEM: row r = lane r, column c = column base + c.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__host__ __device__ inline int digit(uint32_t slice, uint32_t which, uint32_t r, uint32_t k) {
    uint32_t h = (slice * 0x9E3779B1u) ^ (which * 0x85EBCA77u) ^ (r * 0xC2B2AE3Du) ^ (k * 0x27D4EB2Fu);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return (int)(h % 129u) - 64;      // balanced radix-128 digit in [-64, 64]
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// N: MMA N (candidates per CTA tile), S: slices per operand, KC: K bytes resident per slice tile, G: accumulator groups in TMEM
template <int N, int S, int KC, int G>
__global__ void __launch_bounds__(128, 1)
probe_kernel(int iters, int swap_lbo_sbo, int32_t* __restrict__ out, int* __restrict__ err, int group_major = 0) {
    constexpr int M = 128;
    constexpr int A_TILE = M * KC, B_TILE = N * KC;
    constexpr int COLS = (G * N <= 32) ? 32 : (G * N <= 64) ? 64 : (G * N <= 128) ? 128 : (G * N <= 256) ? 256 : 512;
    static_assert(G * N <= 512, "TMEM has 512 columns");
    extern __shared__ __align__(1024) uint8_t smem[];
    int8_t* sA = reinterpret_cast<int8_t*>(smem);
    int8_t* sB = sA + S * A_TILE;
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < S * A_TILE; i += 128) {
        int s = i / A_TILE, o = i % A_TILE, cm = o / 128, w = o % 128;
        int r = (cm / (KC / 16)) * 8 + w / 16, k = (cm % (KC / 16)) * 16 + w % 16;
        sA[i] = (int8_t)digit(s, 0, r, k);
    }
    for (int i = tid; i < S * B_TILE; i += 128) {
        int s = i / B_TILE, o = i % B_TILE, cm = o / 128, w = o % 128;
        int r = (cm / (KC / 16)) * 8 + w / 16, k = (cm % (KC / 16)) * 16 + w % 16;
        sB[i] = (int8_t)digit(s, 1, r, k);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    // instruction descriptor: D = s32, A = B = signed 8 bit, both K-major, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t lbo = 128, sbo = (KC / 16) * 128;
    if (swap_lbo_sbo) { uint32_t t = lbo; lbo = sbo; sbo = t; }
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);

    bool ok = true;
    if (tid == 0) {
        for (int it = 0; it < iters; ++it) {
            if (group_major) {
                // all products of one accumulator back to back: (g+1)*KC/32 MMAs between accumulator switches
#pragma unroll 1
                for (int g = 0; g < S; ++g) {
#pragma unroll 1
                    for (int s = 0; s <= g; ++s) {
#pragma unroll
                        for (int kk = 0; kk < KC / 32; ++kk) {
                            uint64_t da = make_desc(a0 + s * A_TILE + kk * 256, lbo, sbo);
                            uint64_t db = make_desc(b0 + (g - s) * B_TILE + kk * 256, lbo, sbo);
                            mma_i8(tmem_base + (g % G) * N, da, db, idesc, (it > 0 || kk > 0 || s > 0 || g >= G) ? 1u : 0u);
                        }
                    }
                }
            } else
#pragma unroll 1
            for (int s = 0; s < S; ++s) {
#pragma unroll 1
                for (int t = 0; t + s < S; ++t) {
                    const int g = (s + t) % G;
                    // first touch of accumulator g: first iteration, first K step, first pair of the group
                    const bool first_pair = (G == S) ? (s == 0) : (s + t < G && s == 0);
#pragma unroll
                    for (int kk = 0; kk < KC / 32; ++kk) {
                        uint64_t da = make_desc(a0 + s * A_TILE + kk * 256, lbo, sbo);
                        uint64_t db = make_desc(b0 + t * B_TILE + kk * 256, lbo, sbo);
                        mma_i8(tmem_base + g * N, da, db, idesc, (it > 0 || kk > 0 || !first_pair) ? 1u : 0u);
                    }
                }
            }
            commit(&bars[it & 1]);
            if (it > 0) ok = ok && mbar_wait_bounded(&bars[(it - 1) & 1], ((it - 1) >> 1) & 1);
            if (!ok) break;
        }
        if (ok) ok = mbar_wait_bounded(&bars[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
        if (!ok) atomicAdd(err, 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (out != nullptr && blockIdx.x == 0) {
        for (int c0 = 0; c0 < G * N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + (tid & 31)) * (G * N) + c0 + j] = (int32_t)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)COLS) : "memory");
}

template <int N, int S, int KC, int G>
static bool check(int swap, int group_major = 0) {
    constexpr int M = 128;
    size_t smem = (size_t)S * (M + N) * KC;
    CK(cudaFuncSetAttribute(probe_kernel<N, S, KC, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int32_t* d_out; int* d_err;
    CK(cudaMalloc(&d_out, sizeof(int32_t) * M * G * N));
    CK(cudaMemset(d_out, 0xff, sizeof(int32_t) * M * G * N));
    CK(cudaMalloc(&d_err, sizeof(int))); CK(cudaMemset(d_err, 0, sizeof(int)));
    probe_kernel<N, S, KC, G><<<1, 128, smem>>>(1, swap, d_out, d_err, group_major);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> out((size_t)M * G * N);
    int err = 0;
    CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    long bad = 0; int32_t maxabs = 0;
    for (int g = 0; g < G; ++g)
        for (int r = 0; r < M; ++r)
            for (int c = 0; c < N; ++c) {
                long ref = 0;
                for (int s = 0; s <= g && s < S; ++s) {
                    int t = g - s;
                    if (t >= S) continue;
                    for (int k = 0; k < KC; ++k) ref += (long)digit(s, 0, r, k) * digit(t, 1, c, k);
                }
                int32_t got = out[(size_t)r * (G * N) + g * N + c];
                if ((long)got != ref) ++bad;
                if (abs((int)ref) > maxabs) maxabs = abs((int)ref);
            }
    printf("check N=%d S=%d KC=%d swap_lbo_sbo=%d group_major=%d: %ld of %d accumulators differ (max |ref| = %d)%s\n", N, S, KC, swap, group_major, bad, M * G * N, maxabs,
           err ? "  [mbarrier wait timed out]" : "");
    cudaFree(d_out); cudaFree(d_err);
    return bad == 0 && !err;
}

template <int N, int S, int KC, int G>
static void rate(int swap, int sms, int iters, int group_major = 0) {
    constexpr int M = 128;
    size_t smem = (size_t)S * (M + N) * KC;
    CK(cudaFuncSetAttribute(probe_kernel<N, S, KC, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int* d_err; CK(cudaMalloc(&d_err, sizeof(int))); CK(cudaMemset(d_err, 0, sizeof(int)));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        probe_kernel<N, S, KC, G><<<sms, 128, smem>>>(iters, swap, nullptr, d_err, group_major);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    int err = 0; CK(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    const double pairs = S * (S + 1) / 2.0;
    const double macs = (double)sms * iters * pairs * (KC / 32) * M * N * 32.0;
    // the smem fill is part of the launch; iters is chosen so that it is < 1 % of the run
    printf("rate  N=%3d S=%d KC=%3d G=%d %s: %8.3f ms  %7.1f TOP/s int8  -> %6.1f TFLOP/s FP64-equivalent (%d products per FMA)%s\n", N, S, KC, G, group_major ? "group-major" : "slice-major", best,
           2.0 * macs / (best * 1e-3) / 1e12, 2.0 * macs / pairs / (best * 1e-3) / 1e12, (int)pairs, err ? "  [TIMEOUT]" : "");
    cudaFree(d_err);
}

// ---- A operand from TMEM (tcgen05.mma "TS" form): the S slices of one K=32 step of A sit in 8 TMEM columns each
// (lane = row, column j = bytes k=4j..4j+3), next to the S group accumulators: S*N + S*8 <= 512 columns. ----
__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

template <int N, int S>
__global__ void __launch_bounds__(128, 1)
probe_kernel_ts(int iters, int32_t* __restrict__ out, int* __restrict__ err, int group_major = 0) {
    constexpr int M = 128, KC = 32;
    constexpr int B_TILE = N * KC;
    static_assert(S * N + S * 8 <= 512, "TMEM has 512 columns");
    extern __shared__ __align__(1024) uint8_t smem[];
    int8_t* sB = reinterpret_cast<int8_t*>(smem);
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < S * B_TILE; i += 128) {
        int s = i / B_TILE, o = i % B_TILE, cm = o / 128, w = o % 128;
        int r = (cm / (KC / 16)) * 8 + w / 16, k = (cm % (KC / 16)) * 16 + w % 16;
        sB[i] = (int8_t)digit(s, 1, r, k);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t a_col0 = S * N;
    // every thread parks its row of every A slice: 32 bytes = 8 words, byte b of word j = k = 4j + b
    for (int s = 0; s < S; ++s) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) v |= (uint32_t)(uint8_t)(int8_t)digit(s, 0, tid, 4 * j + b) << (8 * b);
            w[j] = v;
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"r"(tmem_base + ((uint32_t)(warp * 32) << 16) + a_col0 + s * 8),
                       "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t b0 = smem_u32(sB);
    bool ok = true;
    if (tid == 0) {
        for (int it = 0; it < iters; ++it) {
            if (group_major) {
#pragma unroll 1
                for (int g = 0; g < S; ++g) {
#pragma unroll 1
                    for (int s = 0; s <= g; ++s) {
                        uint64_t db = make_desc(b0 + (g - s) * B_TILE, 128, (KC / 16) * 128);
                        mma_i8_ts(tmem_base + g * N, tmem_base + a_col0 + s * 8, db, idesc, (it > 0 || s > 0) ? 1u : 0u);
                    }
                }
            } else
#pragma unroll 1
            for (int s = 0; s < S; ++s) {
#pragma unroll 1
                for (int t = 0; t + s < S; ++t) {
                    uint64_t db = make_desc(b0 + t * B_TILE, 128, (KC / 16) * 128);
                    mma_i8_ts(tmem_base + (s + t) * N, tmem_base + a_col0 + s * 8, db, idesc, (it > 0 || s > 0) ? 1u : 0u);
                }
            }
            commit(&bars[it & 1]);
            if (it > 0) ok = ok && mbar_wait_bounded(&bars[(it - 1) & 1], ((it - 1) >> 1) & 1);
            if (!ok) break;
        }
        if (ok) ok = mbar_wait_bounded(&bars[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
        if (!ok) atomicAdd(err, 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (out != nullptr && blockIdx.x == 0) {
        for (int c0 = 0; c0 < S * N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + (tid & 31)) * (S * N) + c0 + j] = (int32_t)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int N, int S>
static bool check_ts() {
    constexpr int M = 128, KC = 32;
    size_t smem = (size_t)S * N * KC;
    CK(cudaFuncSetAttribute(probe_kernel_ts<N, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int32_t* d_out; int* d_err;
    CK(cudaMalloc(&d_out, sizeof(int32_t) * M * S * N));
    CK(cudaMemset(d_out, 0xff, sizeof(int32_t) * M * S * N));
    CK(cudaMalloc(&d_err, sizeof(int))); CK(cudaMemset(d_err, 0, sizeof(int)));
    probe_kernel_ts<N, S><<<1, 128, smem>>>(1, d_out, d_err);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> out((size_t)M * S * N);
    int err = 0;
    CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int g = 0; g < S; ++g)
        for (int r = 0; r < M; ++r)
            for (int c = 0; c < N; ++c) {
                long ref = 0;
                for (int s = 0; s <= g; ++s)
                    for (int k = 0; k < KC; ++k) ref += (long)digit(s, 0, r, k) * digit(g - s, 1, c, k);
                if ((long)out[(size_t)r * (S * N) + g * N + c] != ref) ++bad;
            }
    printf("check TS (A slices in TMEM) N=%d S=%d: %ld of %d accumulators differ%s\n", N, S, bad, M * S * N, err ? "  [mbarrier wait timed out]" : "");
    cudaFree(d_out); cudaFree(d_err);
    return bad == 0 && !err;
}

template <int N, int S>
static void rate_ts(int sms, int iters, int group_major = 0) {
    constexpr int M = 128, KC = 32;
    size_t smem = (size_t)S * N * KC;
    CK(cudaFuncSetAttribute(probe_kernel_ts<N, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int* d_err; CK(cudaMalloc(&d_err, sizeof(int))); CK(cudaMemset(d_err, 0, sizeof(int)));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        probe_kernel_ts<N, S><<<sms, 128, smem>>>(iters, nullptr, d_err, group_major);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    int err = 0; CK(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    const double pairs = S * (S + 1) / 2.0;
    const double macs = (double)sms * iters * pairs * M * N * 32.0;
    printf("rate TS N=%3d S=%d (A in TMEM, B in smem) %s: %8.3f ms  %7.1f TOP/s int8  -> %6.1f TFLOP/s FP64-equivalent%s\n", N, S, group_major ? "group-major" : "slice-major", best,
           2.0 * macs / (best * 1e-3) / 1e12, 2.0 * macs / pairs / (best * 1e-3) / 1e12, err ? "  [TIMEOUT]" : "");
    cudaFree(d_err);
}

// ---- operand feed: every CTA streams `tile` -byte chunks of one buffer (1-D TMA bulk copies, `STAGES` in flight) and does
// nothing else: the aggregate L2 -> shared-memory rate when the buffer is L2-resident (the 59 MB of L^-1 slices at n=4096),
// the HBM rate when it is not. ----
template <int STAGES>
__global__ void __launch_bounds__(128, 1)
feed_kernel(const uint8_t* __restrict__ buf, size_t buf_bytes, uint32_t tile, int tiles_per_cta, int* __restrict__ err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const size_t ntiles = buf_bytes / tile;
        size_t pos = ((size_t)blockIdx.x * 7919u) % ntiles;       // CTAs start at scattered tiles and walk the whole buffer
        bool ok = true;
        for (int i = 0; i < tiles_per_cta + STAGES && ok; ++i) {
            if (i >= STAGES) ok = mbar_wait_bounded(&full[i % STAGES], ((i / STAGES) - 1) & 1);
            if (i < tiles_per_cta && ok) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[i % STAGES])), "r"(tile) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem + (size_t)(i % STAGES) * tile)), "l"(buf + pos * tile), "r"(tile), "r"(smem_u32(&full[i % STAGES])) : "memory");
                pos = (pos + 1 == ntiles) ? 0 : pos + 1;
            }
        }
        if (!ok) atomicAdd(err, 1);
    }
}

static void feed(const char* what, size_t buf_bytes, uint32_t tile, int sms, int tiles_per_cta) {
    constexpr int STAGES = 3;
    uint8_t* d; CK(cudaMalloc(&d, buf_bytes)); CK(cudaMemset(d, 1, buf_bytes));
    int* d_err; CK(cudaMalloc(&d_err, sizeof(int))); CK(cudaMemset(d_err, 0, sizeof(int)));
    CK(cudaFuncSetAttribute(feed_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(STAGES * tile)));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        feed_kernel<STAGES><<<sms, 128, STAGES * tile>>>(d, buf_bytes, tile, tiles_per_cta, d_err);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    int err = 0; CK(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    const double bytes = (double)sms * tiles_per_cta * tile;
    printf("feed  %-40s tile %6u B x %d in flight: %8.3f ms  %7.1f GB/s aggregate = %5.1f B/clk/SM at 1.965 GHz%s\n", what, tile, STAGES, best,
           bytes / (best * 1e-3) / 1e9, bytes / (best * 1e-3) / sms / 1.965e9, err ? "  [TIMEOUT]" : "");
    cudaFree(d); cudaFree(d_err);
}

// ---- the real kernel's stage pattern, issue loop fully unrolled (warp-uniform, one elected lane), in SM cycles:
// MODE 0: SS - 28 slice pairs x 2 K-steps with both operands from shared memory (what sweep_i8_kernel does);
// MODE 1: TS - per K-step the 7 A slices are first copied smem -> TMEM (tcgen05.cp 128x256b into the 56 columns the
//         accumulators leave free) and the 28 MMAs read A from TMEM, B from shared memory. ----
template <int S, int MODE, int KC = 64>
__global__ void __launch_bounds__(128, 1)
probe_kernel_stage(int iters, int32_t* __restrict__ out, int* __restrict__ err, long long* __restrict__ cycles) {
    constexpr int M = 128, N = 64;
    constexpr int A_TILE = M * KC, B_TILE = N * KC;
    static_assert(S * N + (MODE ? S * 8 : 0) <= 512, "TMEM has 512 columns");
    extern __shared__ __align__(1024) uint8_t smem[];
    int8_t* sA = reinterpret_cast<int8_t*>(smem);
    int8_t* sB = sA + S * A_TILE;
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < S * A_TILE; i += 128) {
        int s = i / A_TILE, o = i % A_TILE, cm = o / 128, w = o % 128;
        sA[i] = (int8_t)digit(s, 0, (cm / (KC / 16)) * 8 + w / 16, (cm % (KC / 16)) * 16 + w % 16);
    }
    for (int i = tid; i < S * B_TILE; i += 128) {
        int s = i / B_TILE, o = i % B_TILE, cm = o / 128, w = o % 128;
        sB[i] = (int8_t)digit(s, 1, (cm / (KC / 16)) * 8 + w / 16, (cm % (KC / 16)) * 16 + w % 16);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t a_col0 = S * N;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t da0 = make_desc(smem_u32(sA), 128, (KC / 16) * 128), db0 = make_desc(smem_u32(sB), 128, (KC / 16) * 128);
    if (warp == 1) {
        uint32_t leader;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(leader));
        bool ok = true;
        const long long c0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t acc0 = it > 0 ? 1u : 0u;
            if (leader) {
#pragma unroll
                for (int kk = 0; kk < KC / 32; ++kk) {
                    if (MODE == 1) {
#pragma unroll
                        for (int s = 0; s < S; ++s)
                            asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;"
                                         ::"r"(tmem_base + a_col0 + s * 8), "l"(da0 + (uint64_t)((s * A_TILE + kk * 256) >> 4)) : "memory");
                    }
#pragma unroll
                    for (int s = 0; s < S; ++s)
#pragma unroll
                        for (int t = 0; t + s < S; ++t) {
                            const uint32_t accf = (kk > 0 || s > 0) ? 1u : acc0;
                            if (MODE == 1)
                                mma_i8_ts(tmem_base + (s + t) * N, tmem_base + a_col0 + s * 8, db0 + (uint64_t)((t * B_TILE + kk * 256) >> 4), idesc, accf);
                            else
                                mma_i8(tmem_base + (s + t) * N, da0 + (uint64_t)((s * A_TILE + kk * 256) >> 4),
                                       db0 + (uint64_t)((t * B_TILE + kk * 256) >> 4), idesc, accf);
                        }
                }
                commit(&bars[it & 1]);
            }
            __syncwarp();
            if (it > 0) ok = ok && mbar_wait_bounded(&bars[(it - 1) & 1], ((it - 1) >> 1) & 1);
            if (!ok) break;
        }
        if (ok) ok = mbar_wait_bounded(&bars[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
        const long long c1 = clock64();
        if (!ok && leader) atomicAdd(err, 1);
        if (leader && blockIdx.x == 0 && cycles) *cycles = c1 - c0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (out != nullptr && blockIdx.x == 0) {
        for (int c0 = 0; c0 < S * N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + (tid & 31)) * (S * N) + c0 + j] = (int32_t)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int S, int MODE, int KC = 64>
static void stage_probe(int sms) {
    constexpr int M = 128, N = 64;
    size_t smem = (size_t)S * (M + N) * KC;
    CK(cudaFuncSetAttribute(probe_kernel_stage<S, MODE, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int32_t* d_out; int* d_err; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(int32_t) * M * S * N)); CK(cudaMemset(d_out, 0xff, sizeof(int32_t) * M * S * N));
    CK(cudaMalloc(&d_err, sizeof(int))); CK(cudaMemset(d_err, 0, sizeof(int)));
    CK(cudaMalloc(&d_cyc, sizeof(long long)));
    probe_kernel_stage<S, MODE, KC><<<1, 128, smem>>>(1, d_out, d_err, d_cyc);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> out((size_t)M * S * N);
    CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int g = 0; g < S; ++g)
        for (int r = 0; r < M; ++r)
            for (int c = 0; c < N; ++c) {
                long ref = 0;
                for (int s = 0; s <= g; ++s)
                    for (int k = 0; k < KC; ++k) ref += (long)digit(s, 0, r, k) * digit(g - s, 1, c, k);
                if ((long)out[(size_t)r * (S * N) + g * N + c] != ref) ++bad;
            }
    const int iters = 4000;
    probe_kernel_stage<S, MODE, KC><<<sms, 128, smem>>>(iters, nullptr, d_err, d_cyc);
    CK(cudaDeviceSynchronize());
    long long cyc = 0; int err = 0;
    CK(cudaMemcpy(&cyc, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    const int mmas = S * (S + 1) / 2 * (KC / 32);
    printf("stage %s S=%d KC=%d: %ld of %d accumulators differ; %.0f clk per stage = %.1f clk per MMA (%d MMAs%s)%s\n", MODE ? "TS (A slices copied smem -> TMEM per K-step)" : "SS (both operands from smem)",
           S, KC, bad, M * S * N, (double)cyc / iters, (double)cyc / iters / mmas, mmas, MODE ? " + 14 tcgen05.cp" : "", err ? "  [TIMEOUT]" : "");
    cudaFree(d_out); cudaFree(d_err); cudaFree(d_cyc);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs, cc %d.%d\n", p.name, p.multiProcessorCount, p.major, p.minor);
    bool ok0 = check<64, 7, 64, 7>(0);
    bool ok1 = ok0 ? false : check<64, 7, 64, 7>(1);
    if (!ok0 && !ok1) { printf("no descriptor variant reproduces the integer reference - stopping before the rate runs\n"); return 1; }
    const int swap = ok0 ? 0 : 1;
    const int sms = p.multiProcessorCount;
    stage_probe<7, 0>(sms);
    stage_probe<7, 1>(sms);
    stage_probe<8, 0>(sms);
    stage_probe<7, 0, 32>(sms);      // one K-step per slice pair: would a 4-deep ring of 42 KB stages cost MMA rate?
    stage_probe<7, 0, 128>(sms);     // four K-steps per slice pair (does not fit twice into shared memory)
    return 0;
    rate<64, 7, 64, 7>(swap, sms, 4000);     // the Ozaki shape: 7 group accumulators x 64 columns = 448 TMEM columns
    rate<64, 8, 64, 8>(swap, sms, 4000);     // 8 slices (FP64-equal accuracy): 512 columns
    rate<128, 7, 64, 4>(swap, sms, 2000);    // rate only (groups folded mod 4): what N=128 would give if TMEM were larger
    rate<256, 7, 64, 2>(swap, sms, 1000);    // rate only: N=256
    rate<64, 7, 128, 7>(swap, sms, 2000);    // deeper resident K
    rate<64, 7, 32, 7>(swap, sms, 8000);     // one K step per accumulator visit (what the TS form below does)
    if (check_ts<64, 7>()) rate_ts<64, 7>(sms, 8000);
    if (check<64, 7, 64, 7>(swap, 1)) {
        rate<64, 7, 64, 7>(swap, sms, 4000, 1);
        rate<64, 7, 128, 7>(swap, sms, 2000, 1);
        rate<64, 7, 32, 7>(swap, sms, 8000, 1);
    }
    rate_ts<64, 7>(sms, 8000, 1);
    feed("59 MB buffer (L2-resident)", (size_t)59 << 20, 57344, sms, 4000);
    feed("59 MB buffer (L2-resident)", (size_t)59 << 20, 28672, sms, 8000);
    feed("1 GB buffer (HBM)", (size_t)1 << 30, 57344, sms, 4000);
    return 0;
}
