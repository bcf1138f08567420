// Microbenchmark of the panel kernel's triangular solve X L^T = A (64 rows per CTA, 64 x 64 factor L in shared memory):
// v0 = the production forward substitution (4 lanes per row, one column per step), v1 = blocked by 8 columns with the
// 8 x 8 diagonal leaves inverted up front.  Development aid.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
constexpr int NB = 64;

__global__ void __launch_bounds__(256) trsm_v0(const double* __restrict__ Lg, const double* __restrict__ Ag, double* __restrict__ Xg, long long* cyc) {
    __shared__ double S[NB][NB + 1];
    __shared__ double rdiag[NB];
    const int tid = threadIdx.x;
    for (int e = tid; e < NB * NB; e += 256) S[e / NB][e % NB] = Lg[e];
    __syncthreads();
    const long long t0 = clock64();
    if (tid < NB) rdiag[tid] = 1.0 / S[tid][tid];
    __syncthreads();
    const double* P = Ag + (size_t)blockIdx.x * NB * NB;
    const int row = tid >> 2, q = tid & 3, lane = tid & 31;
    double av[NB / 4], x[NB / 4];
#pragma unroll
    for (int m = 0; m < NB / 4; ++m) { av[m] = P[(size_t)row * NB + 4 * m + q]; x[m] = 0.0; }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        double pp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int m = 0; m < NB / 4; ++m)
            if (4 * m + q < j) pp[m & 3] = fma(x[m], S[j][4 * m + q], pp[m & 3]);
        double part = (pp[0] + pp[1]) + (pp[2] + pp[3]);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        const double aj = __shfl_sync(0xffffffffu, av[j >> 2], (lane & ~3) | (j & 3));
        const double xj = (aj - part) * rdiag[j];
        if ((j & 3) == q) x[j >> 2] = xj;
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    double* O = Xg + (size_t)blockIdx.x * NB * NB;
#pragma unroll
    for (int m = 0; m < NB / 4; ++m) O[(size_t)row * NB + 4 * m + q] = x[m];
}

// blocked: columns in 8 groups of 8.  Thread (row, q): row = tid / 4, lanes q = 0..3 split every dot product 4 ways.
// Step b: t[c] = a[8b + c] - sum_{k < 8b} x[k] L[8b + c][k]  (c = 0..7; independent chains), then x[8b..8b+8) = t * Linv_bb^T
// with the 8 x 8 inverse of the diagonal leaf (computed once, all 8 leaves in parallel).
__global__ void __launch_bounds__(256) trsm_v1(const double* __restrict__ Lg, const double* __restrict__ Ag, double* __restrict__ Xg, long long* cyc) {
    __shared__ double S[NB][NB + 1];
    __shared__ double Inv[8][8][9];          // Inv[b][r][c] = (L_bb^-1)[r][c], lower triangular
    const int tid = threadIdx.x;
    for (int e = tid; e < NB * NB; e += 256) S[e / NB][e % NB] = Lg[e];
    __syncthreads();
    const long long t0 = clock64();
    if (tid < 64) {                           // thread (b, c): column c of the inverse of leaf b by forward substitution
        const int b = tid >> 3, c = tid & 7;
        double xv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) if (k < i) s = fma(-S[8 * b + i][8 * b + k], xv[k], s);
            xv[i] = s / S[8 * b + i][8 * b + i];
            Inv[b][i][c] = (i >= c) ? xv[i] : 0.0;
        }
    }
    __syncthreads();
    const double* P = Ag + (size_t)blockIdx.x * NB * NB;
    const int row = tid >> 2, q = tid & 3;
    // lane q owns columns k with k % 4 == q (as the production kernel): x[m] = X[row][4m + q]
    double av[NB / 4], x[NB / 4];
#pragma unroll
    for (int m = 0; m < NB / 4; ++m) { av[m] = P[(size_t)row * NB + 4 * m + q]; x[m] = 0.0; }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        double t[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double s = 0.0;
#pragma unroll
            for (int m = 0; m < 2 * b; ++m) s = fma(x[m], S[8 * b + c][4 * m + q], s);      // this lane's share of the dot product
            t[c] = s;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            t[c] += __shfl_xor_sync(0xffffffffu, t[c], 1);
            t[c] += __shfl_xor_sync(0xffffffffu, t[c], 2);
        }
        // a values of the 8 columns: column 8b + c lives in lane (c % 4) at m = 2b + c / 4
        double rhs[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) rhs[c] = __shfl_sync(0xffffffffu, av[2 * b + (c >> 2)], (threadIdx.x & 28) | (c & 3)) - t[c];
        // x[8b + r] = sum_{c <= r} rhs[c] * Inv[b][r][c]; this lane keeps r = q and r = q + 4
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = q + 4 * h;
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) s = fma(rhs[c], Inv[b][r][c], s);
            x[2 * b + h] = s;
        }
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    double* O = Xg + (size_t)blockIdx.x * NB * NB;
#pragma unroll
    for (int m = 0; m < NB / 4; ++m) O[(size_t)row * NB + 4 * m + q] = x[m];
}

int main() {
    const int G = 54;
    static double hL[NB * NB], hA[G * NB * NB], x0[G * NB * NB], x1[G * NB * NB];
    for (int i = 0; i < NB; ++i) for (int k = 0; k < NB; ++k) hL[i * NB + k] = (k > i) ? 0.0 : (i == k ? 2.0 + 0.01 * i : 0.3 / (1.0 + i - k));
    for (int e = 0; e < G * NB * NB; ++e) hA[e] = sin(0.37 * e) + 0.1;
    double *L, *A, *X; long long* cyc;
    cudaMalloc(&L, sizeof hL); cudaMalloc(&A, sizeof hA); cudaMalloc(&X, sizeof hA); cudaMalloc(&cyc, 8);
    cudaMemcpy(L, hL, sizeof hL, cudaMemcpyHostToDevice); cudaMemcpy(A, hA, sizeof hA, cudaMemcpyHostToDevice);
    for (int v = 0; v < 2; ++v) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int w = 0; w < 3; ++w) { if (v) trsm_v1<<<G, 256>>>(L, A, X, cyc); else trsm_v0<<<G, 256>>>(L, A, X, cyc); }
        cudaEventRecord(e0);
        for (int r = 0; r < 50; ++r) { if (v) trsm_v1<<<G, 256>>>(L, A, X, cyc); else trsm_v0<<<G, 256>>>(L, A, X, cyc); }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(v ? x1 : x0, X, sizeof hA, cudaMemcpyDeviceToHost);
        printf("trsm v%d: %7.2f us/launch, solve %6lld cycles\n", v, ms * 1e3 / 50, c);
    }
    double md = 0, mx = 0;
    for (int e = 0; e < G * NB * NB; ++e) { md = fmax(md, fabs(x0[e] - x1[e])); mx = fmax(mx, fabs(x0[e])); }
    // residual of v1 against the definition
    double res = 0;
    for (int r = 0; r < NB; ++r) for (int j = 0; j < NB; ++j) { double s = 0; for (int k = 0; k <= j; ++k) s += x1[r * NB + k] * hL[j * NB + k]; res = fmax(res, fabs(s - hA[r * NB + j])); }
    printf("max |v0 - v1| = %g (max |x| = %g), residual of v1 = %g\n", md, mx, res);
    return 0;
}
