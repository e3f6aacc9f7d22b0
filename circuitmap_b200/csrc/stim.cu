// Stimulus-design helpers of the CAVIaR boundary (reference: optimise/caviar.py:35,42).
//
//   cm_caviar_scan_stim   device: number of non-zero entries and the distinct non-zero values of a dense design in ONE
//                         streaming pass (replaces the host-side np.unique / torch.unique the drop-in used to derive
//                         `powers = np.unique(I)[1:]`, caviar.py:42; HBM-bound, s bytes per entry).
//   cm_pack_stim_u8       host: the same scan plus the conversion of a dense float design into uint8 power codes
//                         (code = index into `powers` + 1, 0 = not targeted) with a pool of threads, so that the
//                         reference-facing call uploads N*K bytes instead of 8*N*K (CM_U8 stimulus dtype of cm_caviar_fit).
#include "common.cuh"
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

namespace cm {
namespace stim {

constexpr int SLOTS = CM_CAVIAR_MAX_POWERS + 2;       // one more than the fit accepts, so that "too many" is detectable
constexpr unsigned long long EMPTY = 0xFFFFFFFFFFFFFFFFull;

struct ScanOut {
    unsigned long long nnz;
    unsigned long long slot[SLOTS];
    int overflow;
};

__device__ __forceinline__ void insert_value(ScanOut* o, double v) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    for (int i = 0; i < SLOTS; ++i) {
        const unsigned long long old = atomicCAS(&o->slot[i], EMPTY, bits);
        if (old == EMPTY || old == bits) return;
    }
    o->overflow = 1;
}

template <typename T>
__global__ void __launch_bounds__(256) scan_stim_kernel(const T* __restrict__ x, long long count, ScanOut* o) {
    // every thread remembers the last few distinct values it has already published: designs hold a handful of powers
    double seen[4];
    int nseen = 0;
    unsigned long long nz = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const double v = (double)x[i];
        if (v != 0.0) {
            ++nz;
            bool known = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) known |= (q < nseen) && (seen[q] == v);
            if (!known) {
                insert_value(o, v);
                if (nseen < 4) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (q == nseen) seen[q] = v;
                    ++nseen;
                }
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, off);
    if ((threadIdx.x & 31) == 0 && nz) atomicAdd(&o->nnz, nz);
}

// sparse design (COO triples) -> dense uint8 codes on the device: what a compressive design really is (nnz <= K H entries)
// crosses the bus, the N x K code matrix the index builder streams exists in HBM only
__global__ void __launch_bounds__(256) expand_coo_kernel(const int* __restrict__ neuron, const int* __restrict__ trial,
                                                         const unsigned char* __restrict__ code, long long nnz, int N, int K,
                                                         unsigned char* __restrict__ out, int* status) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int n = neuron[i], k = trial[i];
    if (n < 0 || n >= N || k < 0 || k >= K) { if (status) atomicExch(status, CM_EINVAL); return; }
    out[(size_t)n * K + k] = code[i];
}

// host side ------------------------------------------------------------------------------------------------------
template <typename T>
static void host_distinct(const T* x, long long lo, long long hi, std::vector<double>& vals, std::atomic<int>& too_many,
                          long long& nnz) {
    double last = 0.0;
    long long nz = 0;
    for (long long i = lo; i < hi; ++i) {
        const double v = (double)x[i];
        if (v == 0.0) continue;
        ++nz;
        if (v == last) continue;
        last = v;
        bool known = false;
        for (double u : vals) known |= (u == v) || (u != u && v != v);
        if (!known) {
            if ((int)vals.size() >= SLOTS) { too_many = 1; continue; }
            vals.push_back(v);
        }
    }
    nnz = nz;
}

template <typename T>
static int host_pack(const T* x, long long lo, long long hi, const double* powers, int P, unsigned char* out) {
    int bad = 0;
    double last = 0.0;
    unsigned char lastc = 0;
    for (long long i = lo; i < hi; ++i) {
        const double v = (double)x[i];
        if (v == 0.0) { out[i] = 0; continue; }
        if (v == last) { out[i] = lastc; continue; }
        int pi = -1;
        for (int p = 0; p < P; ++p) if (powers[p] == v) pi = p;
        if (pi < 0) { bad = 1; out[i] = 255; continue; }          // cm_caviar_fit reports code 255 as an invalid entry
        last = v; lastc = (unsigned char)(pi + 1);
        out[i] = lastc;
    }
    return bad;
}

template <typename T>
static int pack_impl(const T* x, long long count, double* powers, int* P_out, int64_t* nnz_out, unsigned char* out,
                     int threads) {
    threads = std::max(1, std::min(threads, 64));
    if (count < (1 << 16)) threads = 1;
    const long long per = (count + threads - 1) / threads;
    std::vector<std::vector<double>> vals(threads);
    std::vector<long long> nnzs(threads, 0);
    std::atomic<int> too_many{0};
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; ++t)
            pool.emplace_back([&, t] { host_distinct(x, std::min(count, t * per), std::min(count, (t + 1) * per), vals[t], too_many, nnzs[t]); });
        host_distinct(x, 0, std::min(count, per), vals[0], too_many, nnzs[0]);
        for (auto& th : pool) th.join();
    }
    std::vector<double> all;
    long long nnz = 0;
    for (int t = 0; t < threads; ++t) {
        nnz += nnzs[t];
        for (double v : vals[t]) {
            bool known = false;
            for (double u : all) known |= (u == v);
            if (v != v) { set_error("cm_pack_stim_u8: the stimulus matrix holds NaN"); return CM_EINVAL; }
            if (!known) all.push_back(v);
        }
    }
    if (nnz < count) all.push_back(0.0);
    std::sort(all.begin(), all.end());
    // powers = np.unique(I)[1:] (caviar.py:42): the sorted distinct values without the smallest one
    const int P = (int)all.size() - 1;
    if (too_many || P > CM_CAVIAR_MAX_POWERS) {
        set_error("cm_pack_stim_u8: more than %d distinct stimulus powers", CM_CAVIAR_MAX_POWERS);
        *P_out = P;
        return CM_EUNSUPPORTED;
    }
    if (all.empty() || all[0] < 0.0) { set_error("cm_pack_stim_u8: the stimulus matrix holds a negative value"); return CM_EINVAL; }
    for (int p = 0; p < P; ++p) powers[p] = all[p + 1];
    *P_out = P < 0 ? 0 : P;
    if (nnz_out) *nnz_out = nnz;
    if (!out) return CM_OK;
    std::vector<int> bad(threads, 0);
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; ++t)
            pool.emplace_back([&, t] { bad[t] = host_pack(x, std::min(count, t * per), std::min(count, (t + 1) * per), powers, P, out); });
        bad[0] = host_pack(x, 0, std::min(count, per), powers, P, out);
        for (auto& th : pool) th.join();
    }
    // a non-zero entry equal to the dropped smallest value (a design without zeros) is kept as code 255: the fit
    // reports it exactly as the float path does ("not among powers")
    return CM_OK;
}

}  // namespace stim
}  // namespace cm

using namespace cm;

extern "C" int cm_caviar_scan_stim(const void* stim_dev, int dtype, int64_t count, void* scratch_dev, int64_t* nnz_out,
                                   double* values_out, int* n_values_out, void* stream) {
    reset_launch_count();
    if (!stim_dev || !scratch_dev || count <= 0 || !nnz_out || !values_out || !n_values_out) {
        set_error("cm_caviar_scan_stim: bad arguments");
        return CM_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    stim::ScanOut init;
    init.nnz = 0; init.overflow = 0;
    for (int i = 0; i < stim::SLOTS; ++i) init.slot[i] = stim::EMPTY;
    stim::ScanOut* o = (stim::ScanOut*)scratch_dev;
    CM_CUDA_CHECK(cudaMemcpyAsync(o, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    int dev = 0, sms = 148;
    CM_CUDA_CHECK(cudaGetDevice(&dev));
    CM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long want = (count + 255) / 256;
    const unsigned grid = (unsigned)std::min<long long>(want, (long long)sms * 8);
    if (dtype == CM_F32) stim::scan_stim_kernel<float><<<grid, 256, 0, st>>>((const float*)stim_dev, count, o);
    else if (dtype == CM_F64) stim::scan_stim_kernel<double><<<grid, 256, 0, st>>>((const double*)stim_dev, count, o);
    else if (dtype == CM_U8) stim::scan_stim_kernel<unsigned char><<<grid, 256, 0, st>>>((const unsigned char*)stim_dev, count, o);
    else { set_error("cm_caviar_scan_stim: bad dtype"); return CM_EINVAL; }
    count_launch();
    CM_CUDA_CHECK(cudaGetLastError());
    stim::ScanOut h;
    CM_CUDA_CHECK(cudaMemcpyAsync(&h, o, sizeof(h), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(cudaStreamSynchronize(st));
    *nnz_out = (int64_t)h.nnz;
    int n = 0;
    for (int i = 0; i < stim::SLOTS; ++i)
        if (h.slot[i] != stim::EMPTY) { double v; memcpy(&v, &h.slot[i], 8); values_out[n++] = v; }
    std::sort(values_out, values_out + n);
    *n_values_out = h.overflow ? stim::SLOTS + 1 : n;
    return CM_OK;
}

extern "C" int cm_expand_stim_coo(const int* neuron_dev, const int* trial_dev, const unsigned char* code_dev, int64_t nnz, int N,
                                  int K, unsigned char* codes_out_dev, int* status_dev, void* stream) {
    reset_launch_count();
    if (!codes_out_dev || N <= 0 || K <= 0 || nnz < 0 || (nnz > 0 && (!neuron_dev || !trial_dev || !code_dev))) {
        set_error("cm_expand_stim_coo: bad arguments");
        return CM_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CM_CUDA_CHECK(cudaMemsetAsync(codes_out_dev, 0, (size_t)N * K, st));
    if (nnz > 0) {
        stim::expand_coo_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(neuron_dev, trial_dev, code_dev, nnz, N, K,
                                                                              codes_out_dev, status_dev);
        count_launch();
    }
    CM_CUDA_CHECK(cudaGetLastError());
    return CM_OK;
}

extern "C" size_t cm_caviar_scan_scratch_bytes(void) { return sizeof(stim::ScanOut); }

extern "C" int cm_pack_stim_u8(const void* stim_host, int dtype, int64_t count, double* powers_out, int* n_powers_out,
                               int64_t* nnz_out, unsigned char* codes_out, int threads) {
    if (!stim_host || count <= 0 || !powers_out || !n_powers_out) { set_error("cm_pack_stim_u8: bad arguments"); return CM_EINVAL; }
    if (dtype == CM_F32) return stim::pack_impl((const float*)stim_host, count, powers_out, n_powers_out, nnz_out, codes_out, threads);
    if (dtype == CM_F64) return stim::pack_impl((const double*)stim_host, count, powers_out, n_powers_out, nnz_out, codes_out, threads);
    set_error("cm_pack_stim_u8: dtype must be CM_F32 or CM_F64");
    return CM_EINVAL;
}
