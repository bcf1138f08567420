// Does the DFMA pipe run concurrently with the DMMA tensor pipe on B200?  (register-resident loops)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// each loop iteration: NM DMMAs (8 independent accumulators) and NF DFMAs (16 independent chains)
template <int NM, int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b) {
    double c[8][2]; double f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r < NM) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            }
            if (r < NF) {
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}

// warp-specialised: even warps DMMA only, odd warps DFMA only
__global__ void __launch_bounds__(256) k_split(double* out, int iters, double a, double b) {
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp & 1) {
        double f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-9 + i;
        for (int it = 0; it < iters * 4; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += f[i];
    } else {
        double c[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    }
    if (s == 123.456) out[0] = s;
}

template <typename F> static double timeit(F launch) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best * 1e-3;
}

template <int NM, int NF> static void run(double* out, int sms) {
    int blocks = sms * 8, iters = 4000; double warps = blocks * 8.0;
    double t = timeit([&] { k_mix<NM, NF><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); });
    double fm = warps * iters * NM * 8 * 512.0, ff = warps * iters * NF * 16 * 32 * 2.0;
    printf("mix NM=%d NF=%d: dmma %.2f + dfma %.2f = %.2f TFLOP/s (%.2f ms)\n", NM, NF, fm / t * 1e-12, ff / t * 1e-12, (fm + ff) / t * 1e-12, t * 1e3);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 8));
    run<8, 0>(out, sms); run<0, 8>(out, sms); run<8, 1>(out, sms); run<8, 2>(out, sms); run<8, 4>(out, sms); run<8, 8>(out, sms); run<4, 8>(out, sms); run<2, 8>(out, sms);
    {
        int blocks = sms * 8, iters = 8000; double warps = blocks * 8.0;
        double t = timeit([&] { k_split<<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); });
        double fm = warps / 2 * iters * 8 * 512.0, ff = warps / 2 * iters * 4 * 16 * 32 * 2.0;
        printf("split warps: dmma %.2f + dfma %.2f = %.2f TFLOP/s (%.2f ms)\n", fm / t * 1e-12, ff / t * 1e-12, (fm + ff) / t * 1e-12, t * 1e3);
    }
    return 0;
}
