// dependent-chain latencies of FP64 ops on B200 (single warp, clock64 timing)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double a, double b) {
    double x = threadIdx.x * 1e-3 + 1.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x = fma(x, a, b);
    }
    long long t1 = clock64();
    double y = x;
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) y = y * a;
    }
    long long t2 = clock64();
    double z = y + 2.0;
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z)); z = r + 1.5; }
    }
    long long t3 = clock64();
    double w = z;
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) w = __shfl_xor_sync(0xffffffffu, w, 1);
    }
    long long t4 = clock64();
    __shared__ double sm[64];
    sm[threadIdx.x] = w; sm[threadIdx.x + 32] = w;
    __syncthreads();
    int idx = threadIdx.x;
    long long t5 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) { idx = (int)sm[idx & 63] & 63; }
    }
    long long t6 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t6 - t5; }
    out[threadIdx.x] = x + y + z + w + idx;
}
__global__ void barlat(long long* cyc) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) { __syncthreads(); }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 256); cudaMallocManaged(&cyc, 64);
    lat<<<1, 32>>>(out, cyc, 1.0000001, 1e-9); cudaDeviceSynchronize();
    barlat<<<1, 256>>>(cyc); cudaDeviceSynchronize();
    printf("DFMA dependent latency  %.1f clk\n", cyc[0] / 16000.0);
    printf("DMUL dependent latency  %.1f clk\n", cyc[1] / 16000.0);
    printf("rsqrt.approx.f64 + DADD %.1f clk\n", cyc[2] / 4000.0);
    printf("SHFL f64 dependent      %.1f clk\n", cyc[3] / 16000.0);
    printf("LDS.64 + cvt dependent  %.1f clk\n", cyc[4] / 16000.0);
    printf("__syncthreads (256 thr) %.1f clk\n", cyc[5] / 1000.0);
    return 0;
}
