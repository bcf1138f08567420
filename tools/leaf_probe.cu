// Microbenchmark of the 64x64 register-resident leaf Cholesky (development aid): where do the ~560 cycles per column step go?
// Variants switch off parts of a step; results are wrong for the stripped variants, only the timing matters.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NB = 64;
template <int MODE>   // 0 full | 1 no rank-1 update | 2 no column capture | 3 __syncwarp instead of __syncthreads | 4 no Newton refinement
__global__ void __launch_bounds__(256) leaf_kernel(const double* __restrict__ A, int lda, double* __restrict__ out, long long* cyc) {
    __shared__ double S[NB][NB + 1];
    __shared__ double colS[2][NB];
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tk = tid & 15;
    double a[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int i = ti + 16 * x, k = tk + 16 * y;
            a[x][y] = (k <= i) ? A[(size_t)i * lda + k] : 0.0;
        }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {
        double fin[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
        for (int jt = 0; jt < 16; ++jt) {
            const int j = jb * 16 + jt;
            double* col = colS[j & 1];
            if (tk == jt) {
#pragma unroll
                for (int x = 0; x < 4; ++x) col[ti + 16 * x] = a[x][jb];
            }
            if (MODE == 3) __syncwarp(); else __syncthreads();
            double dj = col[j];
            if (!(dj > 0.0)) dj = 1.0;
            double rs;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs) : "d"(dj));
            if (MODE != 4) {
                const double hj = 0.5 * dj;
                rs = rs * fma(-hj, rs * rs, 1.5);
                rs = rs * fma(-hj, rs * rs, 1.5);
            }
            double li[4], lk[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) li[x] = col[ti + 16 * x] * rs;
#pragma unroll
            for (int y = 0; y < 4; ++y) lk[y] = col[tk + 16 * y] * rs;
            if (MODE != 1) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y)
                        if (y >= jb) a[x][y] = fma(-li[x], lk[y], a[x][y]);
            } else {
                a[0][jb] += li[0] * lk[0];
            }
            if (MODE == 8) {
                // capture into registers: each thread owns one column per 16-column group, finalised when jt == tk
                double sj = dj * rs;
                sj = fma(0.5 * rs, fma(-sj, sj, dj), sj);
                const bool mine = tk == jt;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int i = ti + 16 * x;
                    const double v = (i > j) ? li[x] : (i == j ? sj : 0.0);
                    fin[x] = mine ? v : fin[x];
                }
            } else if (MODE == 5 || MODE == 6) {
                // branch-free capture: every thread forms the values, only the owners' stores are predicated on
                double sj = dj * rs;
                if (MODE == 5) sj = fma(0.5 * rs, fma(-sj, sj, dj), sj);
                double* dst = &S[ti][j];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int i = ti + 16 * x;
                    const double v = (i > j) ? li[x] : (i == j ? sj : 0.0);
                    if (tk == jt) dst[16 * x * (NB + 1)] = v;
                }
            } else if (MODE == 7) {
                // capture deferred: L[:, j] = col * rs is recomputed from the broadcast column by ONE warp-row of threads later
                if (tid < NB) S[tid][j] = (tid > j) ? col[tid] * rs : (tid == j ? dj * rs : 0.0);
            } else if (MODE != 2 && tk == jt) {
                double sj = dj * rs;
                sj = fma(0.5 * rs, fma(-sj, sj, dj), sj);
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int i = ti + 16 * x;
                    S[i][j] = (i > j) ? li[x] : (i == j ? sj : 0.0);
                }
            }
        }
        if (MODE == 8) {
#pragma unroll
            for (int x = 0; x < 4; ++x) a[x][jb] = fin[x];
        }
    }
    if (MODE == 8) {
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) S[ti + 16 * x][tk + 16 * y] = a[x][y];
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    double s = 0.0;
    for (int e = tid; e < NB * NB; e += 256) s += S[e / NB][e % NB] + a[0][0];
    out[blockIdx.x * 256 + tid] = s;
}
// two columns per barrier: both raw columns are broadcast, every thread redoes the 2x2 pivot block, rank-2 update
__global__ void __launch_bounds__(256) leaf2_kernel(const double* __restrict__ A, int lda, double* __restrict__ out, long long* cyc) {
    __shared__ double S[NB][NB + 1];
    __shared__ double colS[2][2][NB];
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tk = tid & 15;
    double a[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int i = ti + 16 * x, k = tk + 16 * y;
            a[x][y] = (k <= i) ? A[(size_t)i * lda + k] : 0.0;
        }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {
#pragma unroll 1
        for (int jt = 0; jt < 16; jt += 2) {
            const int j = jb * 16 + jt;
            double (*col)[NB] = colS[(jt >> 1) & 1];
            if ((tk & ~1) == jt) {
#pragma unroll
                for (int x = 0; x < 4; ++x) col[tk & 1][ti + 16 * x] = a[x][jb];
            }
            __syncthreads();
            double d0 = col[0][j];
            if (!(d0 > 0.0)) d0 = 1.0;
            double rs0;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs0) : "d"(d0));
            { const double h0 = 0.5 * d0; rs0 = rs0 * fma(-h0, rs0 * rs0, 1.5); rs0 = rs0 * fma(-h0, rs0 * rs0, 1.5); }
            const double l10 = col[0][j + 1] * rs0;
            double d1 = fma(-l10, l10, col[1][j + 1]);
            if (!(d1 > 0.0)) d1 = 1.0;
            double rs1;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs1) : "d"(d1));
            { const double h1 = 0.5 * d1; rs1 = rs1 * fma(-h1, rs1 * rs1, 1.5); rs1 = rs1 * fma(-h1, rs1 * rs1, 1.5); }
            double li0[4], lk0[4], li1[4], lk1[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) { li0[x] = col[0][ti + 16 * x] * rs0; li1[x] = fma(-li0[x], l10, col[1][ti + 16 * x]) * rs1; }
#pragma unroll
            for (int y = 0; y < 4; ++y) { lk0[y] = col[0][tk + 16 * y] * rs0; lk1[y] = fma(-lk0[y], l10, col[1][tk + 16 * y]) * rs1; }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
                    if (y >= jb) a[x][y] = fma(-li1[x], lk1[y], fma(-li0[x], lk0[y], a[x][y]));
            if ((tk & ~1) == jt) {
                const bool second = tk & 1;
                const double dd = second ? d1 : d0, rr = second ? rs1 : rs0;
                double sj = dd * rr;
                sj = fma(0.5 * rr, fma(-sj, sj, dd), sj);
                const int jj = j + (second ? 1 : 0);
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int i = ti + 16 * x;
                    const double lv = second ? li1[x] : li0[x];
                    S[i][jj] = (i > jj) ? lv : (i == jj ? sj : 0.0);
                }
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    double s = 0.0;
    for (int e = tid; e < NB * NB; e += 256) s += S[e / NB][e % NB] + a[0][0];
    out[blockIdx.x * 256 + tid] = s;
}

template <int MODE>
void run(const char* name, const double* A, double* out, long long* cyc, int grid) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) leaf_kernel<MODE><<<grid, 256>>>(A, NB, out, cyc);
    cudaEventRecord(e0);
    for (int r = 0; r < 50; ++r) leaf_kernel<MODE><<<grid, 256>>>(A, NB, out, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s grid %3d: %7.2f us/launch, leaf loop %6lld cycles = %5.1f cycles/step\n", name, grid, ms * 1e3 / 50, c, c / 64.0);
}
int main() {
    double h[NB * NB];
    for (int i = 0; i < NB; ++i) for (int k = 0; k < NB; ++k) h[i * NB + k] = (i == k) ? 70.0 : 1.0 / (1.0 + abs(i - k));
    double *A, *out; long long* cyc;
    cudaMalloc(&A, sizeof h); cudaMalloc(&out, 148 * 256 * 8); cudaMalloc(&cyc, 8);
    cudaMemcpy(A, h, sizeof h, cudaMemcpyHostToDevice);
    for (int grid : {54}) {
        run<0>("full", A, out, cyc, grid);
        run<1>("no rank-1 update", A, out, cyc, grid);
        run<2>("no column capture", A, out, cyc, grid);
        run<3>("__syncwarp instead of __syncthreads", A, out, cyc, grid);
        run<4>("no Newton refinement of rsqrt", A, out, cyc, grid);
        run<5>("branch-free capture", A, out, cyc, grid);
        run<6>("branch-free capture, no Heron step", A, out, cyc, grid);
        run<7>("capture by threads 0..63 from the broadcast column", A, out, cyc, grid);
        run<8>("capture into registers, dump at the end", A, out, cyc, grid);
    }
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int w = 0; w < 3; ++w) leaf2_kernel<<<54, 256>>>(A, NB, out, cyc);
        cudaEventRecord(e0);
        for (int r = 0; r < 50; ++r) leaf2_kernel<<<54, 256>>>(A, NB, out, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s grid %3d: %7.2f us/launch, leaf loop %6lld cycles = %5.1f cycles/column\n", "two columns per barrier (rank-2)", 54, ms * 1e3 / 50, c, c / 64.0);
        // compare the factor with the single-column version
        double* o2; cudaMalloc(&o2, 148 * 256 * 8);
        leaf_kernel<0><<<1, 256>>>(A, NB, out, cyc); leaf2_kernel<<<1, 256>>>(A, NB, o2, cyc);
        double h1[256], h2[256]; cudaMemcpy(h1, out, sizeof h1, cudaMemcpyDeviceToHost); cudaMemcpy(h2, o2, sizeof h2, cudaMemcpyDeviceToHost);
        double md = 0; for (int i = 0; i < 256; ++i) md = fmax(md, fabs(h1[i] - h2[i]));
        printf("checksum difference single vs rank-2: %g (of %g)\n", md, h1[0]);
    }
    return 0;
}
