// FP64 peak micro-benchmark for B200 (sm_100a): DFMA pipe vs DMMA (mma.sync f64) shapes.
// MEASURED_PEAKS.json has no FP64 entry; this supplies the roofline denominator for the
// FP64 sweep kernels. Register-resident operands only: no memory traffic.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int CH>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b) {
    double c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b) {
    double c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static double timeit(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best * 1e-3;
}

int main(int argc, char** argv) {
    int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, 8));
    int iters = 20000;
    for (int wpb = 1; wpb <= 2; ++wpb) {
        int blocks = sms * 4 * wpb, threads = 256;   // 8 warps/block; 32 or 64 warps per SM
        double warps = (double)blocks * threads / 32;
        {   constexpr int CH = 8;
            double t = timeit([&] { k_dfma<CH><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            printf("DFMA   ch=%d blocks=%d: %.2f TFLOP/s (%.3f ms)\n", CH, blocks, warps * 32 * CH * 2.0 * iters / t * 1e-12, t * 1e3); }
        {   constexpr int CH = 8;
            double t = timeit([&] { k_dmma884<CH><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            printf("DMMA m8n8k4   ch=%d blocks=%d: %.2f TFLOP/s (%.3f ms)\n", CH, blocks, warps * CH * 2.0 * 8 * 8 * 4 * iters / t * 1e-12, t * 1e3); }
        {   constexpr int CH = 8;
            double t = timeit([&] { k_dmma1688<CH><<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); }, 5);
            printf("DMMA m16n8k8  ch=%d blocks=%d: %.2f TFLOP/s (%.3f ms)\n", CH, blocks, warps * CH * 2.0 * 16 * 8 * 8 * (iters / 4) / t * 1e-12, t * 1e3); }
        {   constexpr int CH = 8;
            double t = timeit([&] { k_dmma16816<CH><<<blocks, threads>>>(out, iters / 8, 1.0000001, 1e-9); }, 5);
            printf("DMMA m16n8k16 ch=%d blocks=%d: %.2f TFLOP/s (%.3f ms)\n", CH, blocks, warps * CH * 2.0 * 16 * 8 * 16 * (iters / 8) / t * 1e-12, t * 1e3); }
    }
    // sustained: 3 s of the best DMMA shape to see the power-capped clock
    {
        constexpr int CH = 8; int blocks = sms * 8, threads = 256; double warps = (double)blocks * threads / 32;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0;
        for (; n < 60; ++n) k_dmma884<CH><<<blocks, threads>>>(out, iters * 4, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("DMMA m8n8k4 sustained %d launches: %.2f TFLOP/s over %.1f ms\n", n, n * warps * CH * 2.0 * 256 * iters * 4 / (ms * 1e-3) * 1e-12, ms);
        CK(cudaEventRecord(e0));
        for (n = 0; n < 60; ++n) k_dfma<CH><<<blocks, threads>>>(out, iters * 4, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("DFMA sustained %d launches: %.2f TFLOP/s over %.1f ms\n", n, n * warps * 32 * CH * 2.0 * iters * 4 / (ms * 1e-3) * 1e-12, ms);
    }
    return 0;
}
