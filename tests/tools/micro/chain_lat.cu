// Micro-benchmark (diagnostic, not product): dependent-issue latencies that bound one step of the sequential CAVIaR
// chain on B200 -- fp64 FMA, library exp / division, the fast sigmoid of csrc/caviar_fit.inl, warp shuffles, named barriers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_lat chain_lat.cu && ./chain_lat
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>

__device__ __forceinline__ double sigmoid_lib(double x) { return 1.0 / (1.0 + exp(-x)); }

__device__ __forceinline__ double fast_sigmoid(double x) {
    const double t = -x;
    if (!(fabs(t) <= 700.0)) return 1.0 / (1.0 + exp(t));
    const double nf = rint(t * 1.4426950408889634);
    double r = fma(nf, -6.93147180369123816490e-01, t);
    r = fma(nf, -1.90821492927058770002e-10, r);
    const double r2 = r * r;
    const double q0 = fma(r, 1.0, 1.0);
    const double q1 = fma(r, 1.0 / 6, 0.5);
    const double q2 = fma(r, 1.0 / 120, 1.0 / 24);
    const double q3 = fma(r, 1.0 / 5040, 1.0 / 720);
    const double q4 = fma(r, 1.0 / 362880, 1.0 / 40320);
    const double q5 = fma(r, 1.0 / 39916800, 1.0 / 3628800);
    const double q6 = fma(r, 1.0 / 6227020800.0, 1.0 / 479001600);
    const double r4 = r2 * r2;
    const double s0 = fma(q1, r2, q0);
    const double s1 = fma(q3, r2, q2);
    const double s2 = fma(q5, r2, q4);
    const double r8 = r4 * r4;
    const double u0 = fma(s1, r4, s0);
    const double u1 = fma(q6, r4, s2);
    const double p = fma(u1, r8, u0);
    const long long bits = ((long long)((int)nf + 1023)) << 52;
    const double e = p * __longlong_as_double(bits);
    const double d = 1.0 + e;
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double c = fma(-d, y, 1.0);
    y = fma(c, y, y);
    c = fma(-d, y, 1.0);
    y = fma(c, y, y);
    c = fma(-d, y, 1.0);
    y = fma(c, y, y);
    return y;
}

template <int MODE>
__global__ void lat(double* out, long long* cyc, int iters, double seed) {
    double x = seed + threadIdx.x * 1e-3;
    double acc = 0.0;
    __shared__ double sm[512];
    sm[threadIdx.x] = x;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) x = fma(x, 0.999999, 1e-9);                         // dependent DFMA
        if (MODE == 1) x = sigmoid_lib(x) + 0.25;                          // library exp + IEEE divide
        if (MODE == 2) x = fast_sigmoid(x) + 0.25;
        if (MODE == 3) x = x + __shfl_xor_sync(0xffffffffu, x, 1 << (i % 5));   // 64-bit shuffle + DADD
        if (MODE == 4) { asm volatile("bar.sync 1, 128;"); x += 1.0; }
        if (MODE == 5) x = exp(-x) + 0.1;
        if (MODE == 6) x = 1.0 / (1.0 + x);
        if (MODE == 7) { sm[(threadIdx.x * 7 + i) & 127] = x; __syncwarp(); x = sm[(threadIdx.x * 3 + i) & 127] + 1e-9; __syncwarp(); }
        if (MODE == 8) x = x / 3.0000001 + 1.0;
        acc += x;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = acc + x;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
    const char* names[] = {"dependent DFMA", "sigmoid: lib exp + IEEE div", "sigmoid: fast (poly + rcp Newton)", "shfl64 + DADD",
                           "bar.sync 128 threads", "lib exp", "IEEE div + add", "smem st+ld round trip", "IEEE div by const"};
    const int iters = 2000;
    for (int threads : {32, 128}) {
        printf("threads = %d\n", threads);
        for (int m = 0; m < 9; ++m) {
            for (int rep = 0; rep < 2; ++rep) {
                switch (m) {
                    case 0: lat<0><<<1, threads>>>(out, cyc, iters, 0.3); break;
                    case 1: lat<1><<<1, threads>>>(out, cyc, iters, 0.3); break;
                    case 2: lat<2><<<1, threads>>>(out, cyc, iters, 0.3); break;
                    case 3: lat<3><<<1, threads>>>(out, cyc, iters, 0.3); break;
                    case 4: lat<4><<<1, 128>>>(out, cyc, iters, 0.3); break;
                    case 5: lat<5><<<1, threads>>>(out, cyc, iters, 0.3); break;
                    case 6: lat<6><<<1, threads>>>(out, cyc, iters, 0.3); break;
                    case 7: lat<7><<<1, threads>>>(out, cyc, iters, 0.3); break;
                    case 8: lat<8><<<1, threads>>>(out, cyc, iters, 0.3); break;
                }
                cudaDeviceSynchronize();
            }
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("  %-36s %8.1f cycles / iteration\n", names[m], (double)c / iters);
        }
    }
    // accuracy of the fast sigmoid against the library path
    return 0;
}
