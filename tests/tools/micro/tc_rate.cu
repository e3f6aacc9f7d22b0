// Microbenchmark (not product code): cycles per tcgen05.mma kind::tf32 (M=128, K=8) as a function of N, SS mode,
// no-swizzle K-major operands in shared memory.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128) k(long long* cyc, int N, int reps, int kind, int nacc) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, wid = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    const uint32_t fmt = kind == 0 ? 2u : 1u;     // tf32 : bf16
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        uint64_t da = make_desc(smem_u32(smem), 8192, 128);
        const uint64_t db = make_desc(smem_u32(smem + 96 * 1024), (N / 8) * 128, 128);
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (kind == 0) mma_tf32(tb + (r % nacc) * N, da + (r & 63), db, idesc, 1u);
            else mma_bf16(tb + (r % nacc) * N, da + (r & 63), db, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        t1 = clock64();
        cyc[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u));
}
int main() {
    long long* cyc; cudaMalloc(&cyc, 8);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 4096;
    for (int kind = 0; kind < 2; ++kind)
        for (int N : {16, 32, 64, 128, 256}) {
          for (int nacc : {1, 2, 4, 8}) { if (nacc * N > 256) continue;
            k<<<1, 128, smem>>>(cyc, N, reps, kind, nacc);
            cudaError_t e = cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%s M=128 N=%3d nacc=%d: %.1f cycles/MMA (%s)\n", kind == 0 ? "tf32 K=8 " : "bf16 K=16", N, nacc, (double)h / reps, cudaGetErrorString(e)); }
        }
    return 0;
}
