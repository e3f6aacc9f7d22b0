// Microbenchmark (not product code): fp64 DFMA vs DMMA issue rates on one SM and on the whole chip.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double* out, long long* cyc, int iters) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 0.999;
    double f[8], d[8][2];
    for (int i = 0; i < 8; ++i) { f[i] = i; d[i][0] = i; d[i][1] = -i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fma(f[i], b, a);
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(d[i][0], d[i][1], a, b);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += f[i] + d[i][0] + d[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int mode = 0; mode < 3; ++mode) {
            for (int grid : {1, 148}) {
                if (mode == 0) k<0><<<grid, threads>>>(out, cyc, iters);
                if (mode == 1) k<1><<<grid, threads>>>(out, cyc, iters);
                if (mode == 2) k<2><<<grid, threads>>>(out, cyc, iters);
                cudaDeviceSynchronize();
                long long h[148]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
                double c = (double)h[0];
                double dfma = (mode != 1) ? 8.0 * iters * threads : 0;            // thread-FMAs
                double dm = (mode != 0) ? 8.0 * iters * (threads / 32) * 256 : 0;   // FMAs in DMMAs
                printf("threads %4d grid %3d mode %d: %.0f cycles; DFMA %.1f FMA/clk/SM, DMMA %.1f FMA/clk/SM\n", threads, grid, mode, c, dfma / c, dm / c);
            }
        }
    }
    return 0;
}
