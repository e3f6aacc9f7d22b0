// CAVIaR persistent fit kernels, 8-warp variant (two CTAs per SM): second translation unit of csrc/caviar.cu, see
// caviar_common.cuh.  Reference: circuitmap/optimise/caviar.py:20-316, pava.py:9-88.
#include "caviar_common.cuh"

#define CM_NT 256
#define CM_FITNS fit256
namespace cm { namespace cav {
#include "caviar_fit.inl"
} }
#undef CM_NT
#undef CM_FITNS

namespace cm {
namespace cav {

VariantInfo fit256_info() { return VariantInfo{fit256::NT, fit256::NW, fit256::GCT, fit256::FIT_SMEM_BYTES}; }

template <int PT>
static int launch_pt(FitParams& p, int B, int sm_count, int* queue_dev, cudaStream_t st) {
    const int smem_bytes = fit256::FIT_SMEM_BYTES;
    cudaError_t e = cudaFuncSetAttribute(fit256::caviar_fit_kernel<PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return (int)e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fit256::caviar_fit_kernel<PT>, fit256::NT, (size_t)smem_bytes);
    if (e != cudaSuccess) return (int)e;
    const long long wave = (long long)(occ > 0 ? occ : 1) * sm_count;
    p.ct = 1;
    p.queue = queue_dev;
    fit256::caviar_fit_kernel<PT><<<(unsigned)(B < wave ? B : wave), fit256::NT, smem_bytes, st>>>(p);
    return (int)cudaGetLastError();
}

int fit256_launch(FitParams& p, int n_powers, int B, int sm_count, int* queue_dev, cudaStream_t st) {
    return n_powers <= 4 ? launch_pt<4>(p, B, sm_count, queue_dev, st) : launch_pt<PMAX>(p, B, sm_count, queue_dev, st);
}

int fit256_debug(long long* out32, int enable) {
    long long h[32];
    if (cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof(h)) != cudaSuccess) return 1;
    if (out32) for (int i = 0; i < 32; ++i) out32[i] += h[i];
    long long z[32] = {0};
    if (cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)) != cudaSuccess) return 1;
    if (cudaMemcpyToSymbol(g_phase_enable, &enable, sizeof(int)) != cudaSuccess) return 1;
    return 0;
}

}  // namespace cav
}  // namespace cm
