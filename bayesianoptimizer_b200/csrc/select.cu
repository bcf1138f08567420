// Large top-K of a dense score array on the device (`torch.topk(unc, K_big)` on the CPU,
// optimization/Bayesian7.py:681-682, K_big <= 8000).  Order: value descending, index ascending, NaN last -- the same
// order as the fused sweep's top-k (tk_better), so both paths pick identical candidates.
//   1. radix select (8 passes of 8 bits over an order-preserving 64-bit key) -> the K-th largest key T and the number
//      of elements equal to T that still belong to the result;
//   2. compaction: keys above T are appended through an atomic cursor; ties at T are ranked in index order (block
//      counts -> scan -> scatter) so that the lowest indices win deterministically;
//   3. one CTA sorts the K selected (key, index) pairs with a bitonic network in shared memory.
#include "common.cuh"

namespace bo {

constexpr int SEL_THREADS = 256;

struct SelectState {
    unsigned long long prefix;     // selected high bits of the K-th largest key so far
    long long remaining;           // how many elements are still to be taken from the current bucket
    unsigned long long hist[256];
    long long cursor;              // append position of the "above threshold" elements
    long long tie_total;
};

__device__ __forceinline__ unsigned long long order_key(double v) {
    if (!(v == v)) v = -INFINITY;                                     // NaN ranks last
    if (v == 0.0) v = 0.0;                                            // -0.0 and +0.0 compare equal
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);             // larger value <-> larger key
}

__global__ void select_init_kernel(SelectState* st, long long K) {
    if (threadIdx.x < 256) st->hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { st->prefix = 0; st->remaining = K; st->cursor = 0; st->tie_total = 0; }
}

// histogram of digit `shift` among the elements whose higher bits equal the prefix
__global__ void __launch_bounds__(SEL_THREADS) select_hist_kernel(const double* __restrict__ s, long long N, int shift, SelectState* st) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = st->prefix;
    const unsigned long long himask = shift >= 56 ? 0ULL : ~0ULL << (shift + 8);
    for (long long i = (long long)blockIdx.x * SEL_THREADS + threadIdx.x; i < N; i += (long long)gridDim.x * SEL_THREADS) {
        const unsigned long long k = order_key(s[i]);
        if ((k & himask) == (prefix & himask)) atomicAdd(&h[(k >> shift) & 255], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// walk the digits from the top: the bucket in which the cumulative count reaches `remaining` holds the K-th key
__global__ void select_digit_kernel(SelectState* st, int shift) {
    if (threadIdx.x == 0) {
        long long rem = st->remaining;
        int dsel = 0;
        for (int dgt = 255; dgt >= 0; --dgt) {
            const long long c = (long long)st->hist[dgt];
            if (c >= rem) { dsel = dgt; break; }
            rem -= c;
        }
        st->prefix |= (unsigned long long)dsel << shift;
        st->remaining = rem;
    }
    __syncthreads();
    st->hist[threadIdx.x] = 0;
}

// keys above the threshold go out through the cursor; ties are counted per block for the ordered scatter
__global__ void __launch_bounds__(SEL_THREADS) select_compact_kernel(const double* __restrict__ s, long long N, long long first_index,
                                                                     SelectState* st, unsigned long long* __restrict__ okey,
                                                                     long long* __restrict__ oidx, long long* __restrict__ tie_count,
                                                                     long long per_block) {
    __shared__ long long cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const unsigned long long T = st->prefix;
    const long long lo = (long long)blockIdx.x * per_block, hi = min(N, lo + per_block);
    long long mine = 0;
    for (long long i = lo + threadIdx.x; i < hi; i += SEL_THREADS) {
        const unsigned long long k = order_key(s[i]);
        if (k > T) {
            const long long p = atomicAdd((unsigned long long*)&st->cursor, 1ULL);
            okey[p] = k; oidx[p] = first_index + i;
        } else if (k == T) {
            ++mine;
        }
    }
    if (mine) atomicAdd((unsigned long long*)&cnt, (unsigned long long)mine);
    __syncthreads();
    if (threadIdx.x == 0) tie_count[blockIdx.x] = cnt;
}

__global__ void select_scan_kernel(long long* tie_count, int nblocks, SelectState* st) {
    if (threadIdx.x == 0) {                                   // nblocks <= a few thousand: a serial exclusive scan
        long long run = 0;
        for (int b = 0; b < nblocks; ++b) { const long long c = tie_count[b]; tie_count[b] = run; run += c; }
        st->tie_total = run;
    }
}

// ties in index order: the block's base rank + the in-block rank (one warp-synchronous pass per 256 elements)
__global__ void __launch_bounds__(SEL_THREADS) select_ties_kernel(const double* __restrict__ s, long long N, long long first_index,
                                                                  SelectState* st, unsigned long long* __restrict__ okey,
                                                                  long long* __restrict__ oidx, const long long* __restrict__ tie_base,
                                                                  long long per_block, long long K) {
    __shared__ int wcount[SEL_THREADS / 32];
    __shared__ long long run;
    const unsigned long long T = st->prefix;
    const long long need = st->remaining;                    // ties that belong to the result
    const long long above = K - need;                        // == st->cursor after the compaction
    const long long lo = (long long)blockIdx.x * per_block, hi = min(N, lo + per_block);
    if (threadIdx.x == 0) run = tie_base[blockIdx.x];
    __syncthreads();
    if (run >= need) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long i0 = lo; i0 < hi; i0 += SEL_THREADS) {
        const long long i = i0 + threadIdx.x;
        const bool tie = i < hi && order_key(s[i]) == T;
        const unsigned m = __ballot_sync(0xffffffffu, tie);
        if (lane == 0) wcount[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) { if (w < warp) before += wcount[w]; total += wcount[w]; }
        const long long rank = run + before + __popc(m & ((1u << lane) - 1u));
        if (tie && rank < need) { okey[above + rank] = T; oidx[above + rank] = first_index + i; }
        __syncthreads();
        if (threadIdx.x == 0) run += total;
        __syncthreads();
        if (run >= need) return;
    }
}

// bitonic sort of P = 2^m >= K pairs by (key desc, index asc); entries past the selected count are padding
__global__ void __launch_bounds__(1024) select_sort_kernel(const unsigned long long* __restrict__ okey, const long long* __restrict__ oidx,
                                                           long long have, int P, int K, double* __restrict__ vals, long long* __restrict__ idx) {
    extern __shared__ unsigned long long sm[];
    unsigned long long* key = sm;
    long long* id = reinterpret_cast<long long*>(sm + P);
    for (int e = threadIdx.x; e < P; e += 1024) {
        const bool ok = e < have;
        key[e] = ok ? okey[e] : 0ULL;                        // below every real key (the key of -inf is 0x000f...)
        id[e] = ok ? oidx[e] : 0x7fffffffffffffffLL;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < P / 2; t += 1024) {
                const int a = 2 * t - (t & (stride - 1)), b = a + stride;
                const bool down = (a & size) == 0;            // first-before-second in this block
                const bool a_first = key[a] > key[b] || (key[a] == key[b] && id[a] < id[b]);
                if (a_first != down) {
                    const unsigned long long tk = key[a]; key[a] = key[b]; key[b] = tk;
                    const long long ti = id[a]; id[a] = id[b]; id[b] = ti;
                }
            }
            __syncthreads();
        }
    for (int e = threadIdx.x; e < K; e += 1024) {
        if (e < have) {
            const unsigned long long k = key[e];
            const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
            vals[e] = __longlong_as_double((long long)b);
            idx[e] = id[e];
        } else {
            vals[e] = -INFINITY; idx[e] = -1;
        }
    }
}

int topk_scores_impl(bo_handle* h, const double* scores_dev, int64_t N, int64_t first_index, int K, double* vals_dev,
                     int64_t* idx_dev, cudaStream_t st) {
    if (N < 0 || K < 1 || first_index < 0 || (N > 0 && !scores_dev) || !vals_dev || !idx_dev) return fail(h, BO_E_INVALID, "bo_topk_scores: bad argument");
    if (K > BO_MAX_SELECT) return fail(h, BO_E_CAPACITY, "bo_topk_scores: K exceeds BO_MAX_SELECT");
    BO_CUDA(h, cudaSetDevice(h->device));
    const long long Keff = K < N ? K : N;                    // elements actually selected
    int P = 2; while (P < (Keff > 2 ? Keff : 2)) P <<= 1;
    const int nblocks = (int)std::min<long long>((N + 4095) / 4096 > 0 ? (N + 4095) / 4096 : 1, 8LL * h->sm_count);
    const long long per_block = ((N + nblocks - 1) / nblocks + SEL_THREADS - 1) / SEL_THREADS * SEL_THREADS;
    // workspace: state | keys[P] | idx[P] | tie counts[nblocks]
    const size_t need = sizeof(SelectState) + (size_t)P * 16 + (size_t)nblocks * 8 + 256;
    if (need > h->select_bytes) {
        if (h->select_ws) cudaFree(h->select_ws);
        h->select_ws = nullptr; h->select_bytes = 0;
        BO_CUDA(h, cudaMalloc(&h->select_ws, need));
        h->select_bytes = need;
    }
    SelectState* state = reinterpret_cast<SelectState*>(h->select_ws);
    unsigned long long* okey = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(h->select_ws) + ((sizeof(SelectState) + 127) / 128) * 128);
    long long* oidx = reinterpret_cast<long long*>(okey + P);
    long long* ties = oidx + P;
    if (Keff > 0) {
        select_init_kernel<<<1, 256, 0, st>>>(state, Keff);
        BO_LAUNCH_CHECK(h);
        for (int shift = 56; shift >= 0; shift -= 8) {
            select_hist_kernel<<<nblocks, SEL_THREADS, 0, st>>>(scores_dev, N, shift, state);
            BO_LAUNCH_CHECK(h);
            select_digit_kernel<<<1, 256, 0, st>>>(state, shift);
            BO_LAUNCH_CHECK(h);
        }
        select_compact_kernel<<<nblocks, SEL_THREADS, 0, st>>>(scores_dev, N, first_index, state, okey, oidx, ties, per_block);
        BO_LAUNCH_CHECK(h);
        select_scan_kernel<<<1, 32, 0, st>>>(ties, nblocks, state);
        BO_LAUNCH_CHECK(h);
        select_ties_kernel<<<nblocks, SEL_THREADS, 0, st>>>(scores_dev, N, first_index, state, okey, oidx, ties, per_block, Keff);
        BO_LAUNCH_CHECK(h);
    }
    const size_t smem = (size_t)P * 16;
    BO_CUDA(h, cudaFuncSetAttribute(select_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BO_MAX_SELECT * 16)));
    select_sort_kernel<<<1, 1024, smem, st>>>(okey, oidx, Keff, P, K, vals_dev, (long long*)idx_dev);
    BO_LAUNCH_CHECK(h);
    return 0;
}

}  // namespace bo
