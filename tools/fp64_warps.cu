// How many warps / independent accumulators does one SM sub-partition need to fill the DMMA pipe?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
template <int CH>
__global__ void k_dmma(double* out, int iters, double a, double b) {
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
template <int CH> static void run(double* out, int sms, int warps_per_sm) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int iters = 200000 / CH;
    k_dmma<CH><<<sms, warps_per_sm * 32>>>(out, iters, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); k_dmma<CH><<<sms, warps_per_sm * 32>>>(out, iters, 1.0000001, 1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = (double)sms * warps_per_sm * iters * CH * 512.0;
    printf("warps/SM=%2d chains=%2d: %.2f TFLOP/s\n", warps_per_sm, CH, fl / (ms * 1e-3) * 1e-12);
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 8));
    for (int w : {4, 8, 16}) { run<1>(out, sms, w); run<2>(out, sms, w); run<4>(out, sms, w); run<8>(out, sms, w); run<16>(out, sms, w); run<32>(out, sms, w); }
    return 0;
}
