// Micro test (not product code): one 1-D convolution layer (16 -> 16 channels, 32 taps, valid) as an implicit GEMM on
// tcgen05 (kind::tf32, M=128, N=16, K=8 per MMA), operands in shared memory in the no-swizzle K-major canonical layout,
// accumulator in TMEM.  Checks the descriptor encodings used by csrc/nwd_tc.cu against a CPU loop.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CI = 16, CO = 16, KW = 32, LIN = 193, LOUT = 162;
constexpr int G = CI / 4;                 // 4-channel planes
constexpr int TP = 320;                   // positions per plane (padded so that 2 M-tiles + taps stay inside)
constexpr int KSTEPS = KW * CI / 8;       // 64
constexpr int WBLK = CO * 8;              // floats per K-step block of B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;               // descriptor version 1 (Blackwell)
    return d;                             // layout_type = 0: no swizzle ("interleave")
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(128) k(const float* __restrict__ x /*[CI][LIN]*/, const float* __restrict__ w /*[CO][CI][KW]*/,
                                         float* __restrict__ out /*[CO][LOUT]*/) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* A = reinterpret_cast<float*>(smem);                       // [G][TP][4]
    float* B = A + G * TP * 4;                                       // [KSTEPS][2][CO/8][8][4]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < G * TP * 4; i += 128) {
        const int c = i & 3, t = (i >> 2) % TP, g = (i >> 2) / TP;
        A[i] = t < LIN ? to_tf32(x[(4 * g + c) * LIN + t]) : 0.f;
    }
    for (int i = tid; i < KSTEPS * WBLK; i += 128) {
        const int c = i & 3, col = (i >> 2) & 7, cg = (i >> 5) % (CO / 8), h = (i >> 5) / (CO / 8) % 2, s = i / WBLK;
        const int j = s / (CI / 8), q = s % (CI / 8);                 // tap, channel-group pair
        const int co = 8 * cg + col, ci = 8 * q + 4 * h + c;
        B[i] = to_tf32(w[(co * CI + ci) * KW + j]);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(32u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    // instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CO >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (tid == 0) {
        for (int mt = 0; mt < 2; ++mt) {
            for (int s = 0; s < KSTEPS; ++s) {
                const int j = s / (CI / 8), q = s % (CI / 8);
                const uint32_t a_addr = smem_u32(A + ((size_t)(2 * q) * TP + (128 * mt + j)) * 4);
                const uint64_t da = make_desc(a_addr, TP * 16, 128);           // LBO: next 4-channel plane, SBO: next 8 positions
                const uint64_t db = make_desc(smem_u32(B + (size_t)s * WBLK), (CO / 8) * 128, 128);
                mma_tf32(tb + mt * 16, da, db, idesc, s > 0 ? 1u : 0u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int mt = 0; mt < 2; ++mt) {
        uint32_t v[16];
        const uint32_t taddr = tb + mt * 16 + ((uint32_t)(32 * wid) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int t = 128 * mt + 32 * wid + lane;
        if (t < LOUT)
            for (int co = 0; co < CO; ++co) out[co * LOUT + t] = __uint_as_float(v[co]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32u));
}

int main() {
    std::vector<float> x(CI * LIN), w(CO * CI * KW), ref(CO * LOUT), got(CO * LOUT);
    srand(1);
    for (auto& v : x) v = (rand() / (float)RAND_MAX) - 0.3f;
    for (auto& v : w) v = ((rand() / (float)RAND_MAX) - 0.5f) * 0.2f;
    for (int co = 0; co < CO; ++co)
        for (int t = 0; t < LOUT; ++t) {
            double s = 0;
            for (int ci = 0; ci < CI; ++ci)
                for (int j = 0; j < KW; ++j) s += (double)w[(co * CI + ci) * KW + j] * x[ci * LIN + t + j];
            ref[co * LOUT + t] = (float)s;
        }
    float *dx, *dw, *dout;
    cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dw, w.size() * 4); cudaMalloc(&dout, got.size() * 4);
    cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, got.size() * 4);
    const int smem = (G * TP * 4 + KSTEPS * WBLK) * 4 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<1, 128, smem>>>(dx, dw, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (size_t i = 0; i < ref.size(); ++i) { maxerr = fmax(maxerr, fabs(got[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
    printf("max |err| = %.3e (max |ref| = %.3e) -> %s\n", maxerr, maxref, maxerr < 2e-2 * maxref ? "OK" : "MISMATCH");
    printf("sample got %.5f %.5f %.5f ref %.5f %.5f %.5f\n", got[0], got[1], got[LOUT + 5], ref[0], ref[1], ref[LOUT + 5]);
    return 0;
}
