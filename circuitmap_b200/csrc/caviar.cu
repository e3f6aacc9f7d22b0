// CAVIaR (coordinate-ascent variational inference + isotonic regularisation) for sm_100a.
//
// Replaces optimise.caviar and everything it calls (reference circuitmap/optimise/caviar.py:20-316,
// circuitmap/optimise/pava.py:9-88).  Design:
//   * HBM-bound prologue kernels turn the dense inputs into what the loop needs: per-trace trapz / sum of
//     squares (caviar.py:28,30) and a CSR+CSC index of the stimulus design (supp(lam) is a subset of
//     supp(stim), caviar.py:32-34,216), one streaming pass each.
//   * ONE persistent kernel then runs the whole fit -- all `iters` iterations of block_update_mu,
//     update_lam, update_sigma, update_phi, estimate_spont_act_soft_thresh, then reconnect_spont_cells and
//     the final update_phi -- without returning to the host.  One CTA per fit; B fits run concurrently.
//   * fp64 throughout (the reference runs with jax_enable_x64, caviar.py:12), threefry PRNG stream identical
//     to jax.random's (caviar.py:76,196,209-210,304).
#include "caviar_common.cuh"

#define CM_NT 512
#define CM_FITNS fit512
namespace cm { namespace cav {
#include "caviar_fit.inl"
} }
#undef CM_NT
#undef CM_FITNS

namespace cm {
namespace cav {
// ------------------------------------------------------------------------------------------------ prologue kernels
// a1: y = trapz(psc), ss = sum psc^2 per trace (caviar.py:28,30); one warp per trace, coalesced vector loads.
template <typename T>
__global__ void __launch_bounds__(256) psc_stats_kernel(const T* __restrict__ psc, long long ntraces, int Tn,
                                                        const Layout L, char* ws, int K) {
    const int lane = threadIdx.x & 31;
    const long long tr = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tr >= ntraces) return;
    const T* row = psc + (size_t)tr * Tn;
    double s1 = 0.0, s2 = 0.0;
    for (int t = lane; t < Tn; t += 32) {
        const double v = (double)row[t];
        s1 += v; s2 += v * v;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) {
        const int b = (int)(tr / K), k = (int)(tr - (long long)b * K);
        double* y = reinterpret_cast<double*>(ws + (size_t)b * L.stride + L.y);
        double* ss = reinterpret_cast<double*>(ws + (size_t)b * L.stride + L.ss);
        y[k] = s1 - 0.5 * ((double)row[0] + (double)row[Tn - 1]);
        ss[k] = s2;
        reinterpret_cast<int*>(ws + (size_t)b * L.stride + L.colfill)[k] = 0;      // column counters of csr_count_kernel
    }
}

__global__ void copy_stats_kernel(const double* __restrict__ yin, const double* __restrict__ ssin, const Layout L,
                                  char* ws, int K, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int b = (int)(i / K), k = (int)(i - (long long)b * K);
    reinterpret_cast<double*>(ws + (size_t)b * L.stride + L.y)[k] = yin[i];
    reinterpret_cast<double*>(ws + (size_t)b * L.stride + L.ss)[k] = ssin[i];
    reinterpret_cast<int*>(ws + (size_t)b * L.stride + L.colfill)[k] = 0;
}

struct PowerTable { double v[PMAX]; int P; };

__device__ __forceinline__ int power_index(const PowerTable& pt, double v) {
    for (int p = 0; p < pt.P; ++p) if (pt.v[p] == v) return p;
    return -1;
}

// One entry of the dense design: 0 = not targeted, 1 = targeted with power index pi, -1 = invalid.
// Floating-point designs hold the laser power itself (README.md:26); CM_U8 designs hold the code pi + 1.
template <typename T>
__device__ __forceinline__ int stim_class(const PowerTable& pt, T raw, int& pi) {
    const double v = (double)raw;
    if (v > 0.0) { pi = power_index(pt, v); return pi < 0 ? -1 : 1; }
    return (v < 0.0 || v != v) ? -1 : 0;
}
template <>
__device__ __forceinline__ int stim_class<unsigned char>(const PowerTable& pt, unsigned char raw, int& pi) {
    if (raw == 0) return 0;
    pi = (int)raw - 1;
    return pi < pt.P ? 1 : -1;
}

// pass 1 over the dense design: per-row and per-column counts.  Only trials that pass the lam_mask
// (sum psc^2 > y_xcorr_thresh, caviar.py:30) enter the CSR/CSC index -- lam is identically zero on the others
// (caviar.py:34,216); the per-power trial counts (spike-rate denominators, caviar.py:183) count every trial.
template <typename T>
__global__ void __launch_bounds__(256) csr_count_kernel(const T* __restrict__ stim, int N, int K, const Layout L,
                                                        char* ws, const PowerTable pt, int* status, double thresh) {
    const int n = blockIdx.x, b = blockIdx.y;
    char* base = ws + (size_t)b * L.stride;
    int* row_ptr = reinterpret_cast<int*>(base + L.row_ptr);
    int* colcnt = reinterpret_cast<int*>(base + L.colfill);
    int* cntp = reinterpret_cast<int*>(base + L.cntp);
    int* nmask = reinterpret_cast<int*>(base + L.nmask);
    const double* ss = reinterpret_cast<const double*>(base + L.ss);
    __shared__ int cs[PMAX + 1], cm[PMAX];
    if (threadIdx.x <= PMAX) cs[threadIdx.x] = 0;
    if (threadIdx.x < PMAX) cm[threadIdx.x] = 0;
    __syncthreads();
    const T* row = stim + ((size_t)b * N + n) * K;
    int bad = 0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        int pi = 0;
        const int cls = stim_class<T>(pt, row[k], pi);
        if (cls > 0) {
            atomicAdd(&cs[pi], 1);
            if (ss[k] > thresh) { atomicAdd(&cs[PMAX], 1); atomicAdd(&colcnt[k], 1); }
            else atomicAdd(&cm[pi], 1);
        } else if (cls < 0) bad = 1;
    }
    if (bad) atomicExch(&status[b], CM_EINVAL);
    __syncthreads();
    if (threadIdx.x < PMAX) { cntp[n * PMAX + threadIdx.x] = cs[threadIdx.x]; nmask[n * PMAX + threadIdx.x] = cm[threadIdx.x]; }
    if (threadIdx.x == 0) row_ptr[n + 1] = cs[PMAX];      // counts; scanned next
}

// exclusive scans of the row / column counts (one CTA per fit)
__global__ void __launch_bounds__(1024) scan_kernel(int N, int K, const Layout L, char* ws, long long nnz_cap,
                                                    int* status) {
    const int b = blockIdx.x;
    char* base = ws + (size_t)b * L.stride;
    int* row_ptr = reinterpret_cast<int*>(base + L.row_ptr);
    int* col_ptr = reinterpret_cast<int*>(base + L.col_ptr);
    int* colfill = reinterpret_cast<int*>(base + L.colfill);
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x < 64) reinterpret_cast<int*>(base + L.job)[threadIdx.x] = 0;     // job board of the helper CTAs
    for (int pass = 0; pass < 2; ++pass) {
        const int n = pass == 0 ? N : K;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += 1024) {
            const int i = i0 + threadIdx.x;
            int v = 0;
            if (i < n) v = pass == 0 ? row_ptr[i + 1] : colfill[i];
            int inc = v;
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            if (lane == 31) wsum[wid] = inc;
            __syncthreads();
            int off = carry;
            for (int w = 0; w < wid; ++w) off += wsum[w];
            __syncthreads();
            if (i < n) {
                if (pass == 0) row_ptr[i + 1] = off + inc;         // inclusive -> row_ptr[i+1]
                else { col_ptr[i + 1] = off + inc; colfill[i] = 0; }
            }
            if (threadIdx.x == 1023) carry = off + inc;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            if (pass == 0) { row_ptr[0] = 0; if ((long long)carry > nnz_cap) atomicExch(&status[b], CM_EWORKSPACE); }
            else col_ptr[0] = 0;
        }
        __syncthreads();
    }
}

// pass 2 over the dense design: ordered CSR fill + unordered CSC scatter
template <typename T>
__global__ void __launch_bounds__(256) csr_fill_kernel(const T* __restrict__ stim, int N, int K, const Layout L,
                                                       char* ws, const PowerTable pt, const int* status, double thresh) {
    const int n = blockIdx.x, b = blockIdx.y;
    if (status[b] != 0) return;
    char* base = ws + (size_t)b * L.stride;
    const int* row_ptr = reinterpret_cast<const int*>(base + L.row_ptr);
    const int* col_ptr = reinterpret_cast<const int*>(base + L.col_ptr);
    int* colfill = reinterpret_cast<int*>(base + L.colfill);
    int* col_k = reinterpret_cast<int*>(base + L.col_k);
    int* csc_row = reinterpret_cast<int*>(base + L.csc_row);
    int* csc_pos = reinterpret_cast<int*>(base + L.csc_pos);
    unsigned char* pw = reinterpret_cast<unsigned char*>(base + L.pw);
    int* colpw = reinterpret_cast<int*>(base + L.colpw);
    const double* ss = reinterpret_cast<const double*>(base + L.ss);
    // Ordered compaction of the row, 1024 trials per round (four coalesced sub-chunks of 256): one barrier pair per round
    // instead of three per 256 trials -- the kernel is bound by barrier latency, not by the 8 bytes per trial it reads.
    constexpr int SUB = 4;
    __shared__ int wcnt[SUB][8];
    int run = row_ptr[n];                              // next free CSR slot of the row (identical in every thread)
    const T* row = stim + ((size_t)b * N + n) * K;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k0 = 0; k0 < K; k0 += SUB * 256) {
        T v[SUB];
        int pis[SUB];
        unsigned bal[SUB];
#pragma unroll
        for (int j = 0; j < SUB; ++j) {
            const int k = k0 + j * 256 + threadIdx.x;
            v[j] = (k < K) ? row[k] : (T)0;
        }
#pragma unroll
        for (int j = 0; j < SUB; ++j) {
            const int k = k0 + j * 256 + threadIdx.x;
            pis[j] = 0;
            const bool f = stim_class<T>(pt, v[j], pis[j]) > 0 && ss[k < K ? k : 0] > thresh;
            bal[j] = __ballot_sync(0xffffffffu, f);
            if (lane == 0) wcnt[j][wid] = __popc(bal[j]);
        }
        __syncthreads();
        int off = run;
#pragma unroll
        for (int j = 0; j < SUB; ++j) {
            int mine = off;
            for (int w = 0; w < 8; ++w) { mine += (w < wid) ? wcnt[j][w] : 0; off += wcnt[j][w]; }
            if (bal[j] & (1u << lane)) {
                const int k = k0 + j * 256 + threadIdx.x;
                const int jj = mine + __popc(bal[j] & ((1u << lane) - 1u));
                col_k[jj] = k;
                const int pi = pis[j];
                pw[jj] = (unsigned char)pi;
                colpw[jj] = k | (pi << 27);           // packed (trial, power) for the sweep
                const int slot = atomicAdd(&colfill[k], 1);
                csc_row[col_ptr[k] + slot] = n;
                csc_pos[col_ptr[k] + slot] = jj;
            }
        }
        run = off;
        __syncthreads();                              // wcnt is rewritten in the next round
    }
}

// sort every column list by neuron index (deterministic summation order for pred / Gram)
__global__ void csc_sort_kernel(int K, const Layout L, char* ws, const int* status, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int b = (int)(i / K), k = (int)(i - (long long)b * K);
    if (status[b] != 0) return;
    char* base = ws + (size_t)b * L.stride;
    const int* col_ptr = reinterpret_cast<const int*>(base + L.col_ptr);
    int* csc_row = reinterpret_cast<int*>(base + L.csc_row);
    int* csc_pos = reinterpret_cast<int*>(base + L.csc_pos);
    const int s = col_ptr[k], e = col_ptr[k + 1];
    for (int a = s + 1; a < e; ++a) {
        const int r = csc_row[a], q = csc_pos[a];
        int t = a - 1;
        while (t >= s && csc_row[t] > r) { csc_row[t + 1] = csc_row[t]; csc_pos[t + 1] = csc_pos[t]; --t; }
        csc_row[t + 1] = r; csc_pos[t + 1] = q;
    }
}

// scatter the sparse posterior back into the dense N x K array the reference returns (memset to 0 beforehand)
__global__ void __launch_bounds__(128) densify_kernel(const Layout L, char* ws, int N, int K, double* __restrict__ out,
                                                      int iters_hist, const int* status) {
    const int n = blockIdx.x, b = blockIdx.z;
    const int h = blockIdx.y;                     // history slot (0 when densifying the final state)
    if (status[b] != 0) return;
    char* base = ws + (size_t)b * L.stride;
    const int* row_ptr = reinterpret_cast<const int*>(base + L.row_ptr);
    const int* col_k = reinterpret_cast<const int*>(base + L.col_k);
    const int nnz = row_ptr[N];
    const double* src = iters_hist ? reinterpret_cast<const double*>(base + L.lamhist) + (size_t)h * nnz
                                   : reinterpret_cast<const double*>(base + L.lam);
    double* dst = out + (((size_t)b * (iters_hist ? iters_hist : 1) + h) * N + n) * (size_t)K;
    for (int j = row_ptr[n] + threadIdx.x; j < row_ptr[n + 1]; j += blockDim.x) dst[col_k[j]] = src[j];
}

}  // namespace cav
}  // namespace cm

using namespace cm;
using namespace cm::cav;

extern "C" int cm_caviar_debug_phase_cycles(long long* out, int n, int enable) {
    if (out && n > 0) {
        long long h[32];
        CM_CUDA_CHECK(cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof(h)));
        for (int i = 0; i < n && i < 32; ++i) out[i] = h[i];
    }
    long long z[32] = {0};
    CM_CUDA_CHECK(cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)));
    CM_CUDA_CHECK(cudaMemcpyToSymbol(g_phase_enable, &enable, sizeof(int)));
    {                                                          // the 8-warp unit keeps its own counters
        long long h2[32] = {0};
        if (fit256_debug(h2, enable) != 0) { set_error("cm_caviar_debug_phase_cycles: fit256 unit failed"); return CM_ECUDA; }
        if (out) for (int i = 0; i < n && i < 32; ++i) out[i] += h2[i];
    }
    return CM_OK;
}

// CTA variant of the persistent kernel: 16-warp CTAs (one per SM) unless the batch has at least two fits per SM
static bool use_small_cta(int B, int forced) {
    if (forced == 256) return true;
    if (forced == 512) return false;
    if (const char* f = getenv("CM_CAVIAR_CTA")) {            // diagnostics: force a variant ("256" / "512")
        if (!strcmp(f, "256")) return true;
        if (!strcmp(f, "512")) return false;
    }
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
    // One resident wave holds `sms` 16-warp CTAs or 2 * sms 8-warp CTAs.  A batch that does not fit one wave of the big
    // variant runs the small one: SMs holding two fits overlap their latency-bound phases (measured 1.17x at 2 * sms
    // fits), and SMs < B <= 2 * sms fits finish in one wave instead of two.
    return B > sms;
}

// ---- uint8 power codes: one WARP per row, W 32-bit words (4 W codes) per lane and load ----
// The generic kernels above walk a row with one CTA, one element per thread and two block barriers per 1024 trials: for
// byte codes that is 256-byte requests and ~10 us of barrier latency per row (12 ms per 296 C3 designs, 13 x the time the
// 3 GB take to stream).  Here a warp streams its row with 16-byte loads (99 % of the codes are zero: a lane looks at a
// word only when it is non-zero), counts in warp-private shared memory, and the ordered fill needs one warp scan per
// 512 trials instead of block barriers.  Same outputs as the generic kernels (CSR ordered by trial, CSC lists completed
// by csc_sort_kernel).
template <int W> struct CodeVec;
template <> struct CodeVec<4> { using type = uint4; };
template <> struct CodeVec<2> { using type = uint2; };
template <> struct CodeVec<1> { using type = unsigned; };
template <int W>
__device__ __forceinline__ void load_codes(const unsigned char* p, unsigned (&w)[W]) {
    const typename CodeVec<W>::type q = *reinterpret_cast<const typename CodeVec<W>::type*>(p);
    const unsigned* qq = reinterpret_cast<const unsigned*>(&q);
#pragma unroll
    for (int i = 0; i < W; ++i) w[i] = qq[i];
}
template <int W>
__global__ void __launch_bounds__(256) csr_count_u8_kernel(const unsigned char* __restrict__ stim, int N, int K, const Layout L,
                                                           char* ws, const PowerTable pt, int* status, double thresh) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + wid, b = blockIdx.y;
    __shared__ int cs[8][PMAX + 1], cm[8][PMAX];
    for (int i = threadIdx.x; i < 8 * (PMAX + 1); i += 256) (&cs[0][0])[i] = 0;
    for (int i = threadIdx.x; i < 8 * PMAX; i += 256) (&cm[0][0])[i] = 0;
    __syncthreads();
    if (n >= N) return;
    char* base = ws + (size_t)b * L.stride;
    int* row_ptr = reinterpret_cast<int*>(base + L.row_ptr);
    int* colcnt = reinterpret_cast<int*>(base + L.colfill);
    int* cntp = reinterpret_cast<int*>(base + L.cntp);
    int* nmask = reinterpret_cast<int*>(base + L.nmask);
    const double* ss = reinterpret_cast<const double*>(base + L.ss);
    const unsigned char* row = stim + ((size_t)b * N + n) * K;
    const int nvec = K / (4 * W);
    int bad = 0;
    for (int v = lane; v < nvec; v += 32) {
        unsigned w[W];
        load_codes<W>(row + (size_t)v * 4 * W, w);
#pragma unroll
        for (int i = 0; i < W; ++i) {
            unsigned m = __vcmpne4(w[i], 0u) & 0x01010101u;       // one bit per non-zero code
            while (m) {
                const int j = (__ffs(m) - 1) >> 3;
                m &= m - 1;
                const int k = (v * W + i) * 4 + j, pi = (int)((w[i] >> (8 * j)) & 0xffu) - 1;
                if (pi < pt.P) {
                    atomicAdd(&cs[wid][pi], 1);
                    if (ss[k] > thresh) { atomicAdd(&cs[wid][PMAX], 1); atomicAdd(&colcnt[k], 1); }
                    else atomicAdd(&cm[wid][pi], 1);
                } else bad = 1;
            }
        }
    }
    if (bad) atomicExch(&status[b], CM_EINVAL);
    __syncwarp();
    if (lane < PMAX) { cntp[n * PMAX + lane] = cs[wid][lane]; nmask[n * PMAX + lane] = cm[wid][lane]; }
    if (lane == 0) row_ptr[n + 1] = cs[wid][PMAX];              // counts; scanned next
}
template <int W>
__global__ void __launch_bounds__(256) csr_fill_u8_kernel(const unsigned char* __restrict__ stim, int N, int K, const Layout L,
                                                          char* ws, const PowerTable pt, const int* status, double thresh) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + wid, b = blockIdx.y;
    if (n >= N || status[b] != 0) return;
    char* base = ws + (size_t)b * L.stride;
    const int* row_ptr = reinterpret_cast<const int*>(base + L.row_ptr);
    const int* col_ptr = reinterpret_cast<const int*>(base + L.col_ptr);
    int* colfill = reinterpret_cast<int*>(base + L.colfill);
    int* col_k = reinterpret_cast<int*>(base + L.col_k);
    int* csc_row = reinterpret_cast<int*>(base + L.csc_row);
    int* csc_pos = reinterpret_cast<int*>(base + L.csc_pos);
    unsigned char* pw = reinterpret_cast<unsigned char*>(base + L.pw);
    int* colpw = reinterpret_cast<int*>(base + L.colpw);
    const double* ss = reinterpret_cast<const double*>(base + L.ss);
    const unsigned char* row = stim + ((size_t)b * N + n) * K;
    const int nvec = K / (4 * W);
    int run = row_ptr[n];                                   // next free CSR slot of the row (identical in every lane)
    for (int v0 = 0; v0 < nvec; v0 += 32) {
        const int v = v0 + lane;
        unsigned w[W];
#pragma unroll
        for (int i = 0; i < W; ++i) w[i] = 0u;
        if (v < nvec) load_codes<W>(row + (size_t)v * 4 * W, w);
        unsigned valid = 0;                                 // bit 4 i + j: code j of word i enters the index
#pragma unroll
        for (int i = 0; i < W; ++i) {
            unsigned m = __vcmpne4(w[i], 0u) & 0x01010101u;
            while (m) {
                const int j = (__ffs(m) - 1) >> 3;
                m &= m - 1;
                const int k = (v * W + i) * 4 + j, pi = (int)((w[i] >> (8 * j)) & 0xffu) - 1;
                if (pi < pt.P && ss[k] > thresh) valid |= 1u << (4 * i + j);
            }
        }
        const int cnt = __popc(valid);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        int jj = run + inc - cnt;                           // first slot of this lane's entries (ascending trials)
        while (valid) {
            const int bit = __ffs(valid) - 1;
            valid &= valid - 1;
            const int i = bit >> 2, j = bit & 3;
            const int k = (v * W + i) * 4 + j, pi = (int)((w[i] >> (8 * j)) & 0xffu) - 1;
            col_k[jj] = k;
            pw[jj] = (unsigned char)pi;
            colpw[jj] = k | (pi << 27);                     // packed (trial, power) for the sweep
            const int slot = atomicAdd(&colfill[k], 1);
            csc_row[col_ptr[k] + slot] = n;
            csc_pos[col_ptr[k] + slot] = jj;
            ++jj;
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
}

extern "C" size_t cm_caviar_workspace_bytes(int B, int N, int K, int64_t nnz_cap, int save_histories) {
    if (B <= 0 || N <= 0 || K <= 0 || nnz_cap < 0) return 0;
    // sized for the 16-warp layout (256-column tiles, 16 growbufs), which bounds the 8-warp one
    const Layout L = make_layout(N, K, nnz_cap, save_histories > 0 ? save_histories : 0, save_histories > 0, 256, 16);
    return L.stride * (size_t)B + (size_t)B * 8 + 512;      // + device copy of the seeds + fit queue counter
}

template <int W>
static void run_csr_u8(const cm_caviar_args* a, const Layout& L, char* ws, const PowerTable& pt, cudaStream_t st) {
    dim3 grid((a->N + 7) / 8, a->B);
    const unsigned char* stim = (const unsigned char*)a->stim_dev;
    csr_count_u8_kernel<W><<<grid, 256, 0, st>>>(stim, a->N, a->K, L, ws, pt, a->status_dev, a->opt.y_xcorr_thresh);
    scan_kernel<<<a->B, 1024, 0, st>>>(a->N, a->K, L, ws, (long long)a->nnz_cap, a->status_dev);
    csr_fill_u8_kernel<W><<<grid, 256, 0, st>>>(stim, a->N, a->K, L, ws, pt, a->status_dev, a->opt.y_xcorr_thresh);
}
template <typename TS>
static int run_csr(const cm_caviar_args* a, const Layout& L, char* ws, const PowerTable& pt, cudaStream_t st) {
    dim3 grid(a->N, a->B);
    // byte codes whose rows are 4 / 8 / 16-byte aligned take the warp-per-row kernels
    const int walign = (sizeof(TS) == 1 && a->K % 4 == 0 && ((uintptr_t)a->stim_dev & 15) == 0)
                           ? (a->K % 16 == 0 ? 4 : (a->K % 8 == 0 ? 2 : 1)) : 0;
    if (walign == 4) run_csr_u8<4>(a, L, ws, pt, st);
    else if (walign == 2) run_csr_u8<2>(a, L, ws, pt, st);
    else if (walign == 1) run_csr_u8<1>(a, L, ws, pt, st);
    else {
    csr_count_kernel<TS><<<grid, 256, 0, st>>>((const TS*)a->stim_dev, a->N, a->K, L, ws, pt, a->status_dev,
                                               a->opt.y_xcorr_thresh);
    scan_kernel<<<a->B, 1024, 0, st>>>(a->N, a->K, L, ws, (long long)a->nnz_cap, a->status_dev);
    csr_fill_kernel<TS><<<grid, 256, 0, st>>>((const TS*)a->stim_dev, a->N, a->K, L, ws, pt, a->status_dev,
                                              a->opt.y_xcorr_thresh);
    }
    const long long total = (long long)a->B * a->K;
    csc_sort_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a->K, L, ws, a->status_dev, total);
    count_launch(4);
    CM_CUDA_CHECK(cudaGetLastError());
    return CM_OK;
}

extern "C" int cm_caviar_fit(const cm_caviar_args* a, void* stream) {
    reset_launch_count();
    if (!a) { set_error("cm_caviar_fit: null args"); return CM_EINVAL; }
    if (a->B <= 0 || a->N <= 0 || a->K <= 0) { set_error("cm_caviar_fit: bad shape B=%d N=%d K=%d", a->B, a->N, a->K); return CM_ESHAPE; }
    if (a->n_powers < 1 || a->n_powers > PMAX) {
        set_error("cm_caviar_fit: %d distinct powers unsupported (1..%d)", a->n_powers, PMAX);
        return CM_EUNSUPPORTED;
    }
    if (!a->stim_dev || !a->powers || !a->seeds || !a->mu0_dev || !a->beta0_dev || !a->phi0_dev || !a->phi_cov0_dev ||
        !a->shape0 || !a->rate0 || !a->mu_dev || !a->beta_dev || !a->shape_dev || !a->rate_dev || !a->phi_dev ||
        !a->phi_cov_dev || !a->z_dev || !a->workspace_dev || !a->status_dev) {
        set_error("cm_caviar_fit: null pointer among required arguments");
        return CM_EINVAL;
    }
    if (!a->psc_dev && !(a->y_dev && a->ss_dev)) { set_error("cm_caviar_fit: need psc_dev or (y_dev, ss_dev)"); return CM_EINVAL; }
    if (a->opt.iters < 0 || a->opt.num_mc_samples < 1) { set_error("cm_caviar_fit: bad iters/num_mc_samples"); return CM_EINVAL; }
    const bool want_lamhist = a->opt.save_histories && a->lam_hist_dev;
    if (a->cta_variant != 0 && a->cta_variant != 256 && a->cta_variant != 512) { set_error("cm_caviar_fit: cta_variant must be 0, 256 or 512"); return CM_EINVAL; }
    const bool small_cta = use_small_cta(a->B, a->cta_variant);
    const VariantInfo v256 = fit256_info();
    const Layout L = small_cta ? make_layout(a->N, a->K, a->nnz_cap, a->opt.iters, want_lamhist, v256.GCT, v256.NW)
                               : make_layout(a->N, a->K, a->nnz_cap, a->opt.iters, want_lamhist, fit512::GCT, fit512::NW);
    const size_t need = L.stride * (size_t)a->B + (size_t)a->B * 8 + 512;
    if (a->workspace_bytes < need) {
        set_error("cm_caviar_fit: workspace %zu < required %zu bytes", a->workspace_bytes, need);
        return CM_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)a->workspace_dev;
    unsigned long long* seeds_dev = (unsigned long long*)(ws + L.stride * (size_t)a->B);
    seeds_dev = (unsigned long long*)(((uintptr_t)seeds_dev + 255) & ~(uintptr_t)255);
    CM_CUDA_CHECK(cudaMemcpyAsync(seeds_dev, a->seeds, (size_t)a->B * 8, cudaMemcpyHostToDevice, st));
    int* queue_dev = (int*)(((uintptr_t)(seeds_dev + a->B) + 63) & ~(uintptr_t)63);
    CM_CUDA_CHECK(cudaMemsetAsync(a->status_dev, 0, (size_t)a->B * sizeof(int), st));
    CM_CUDA_CHECK(cudaMemsetAsync(queue_dev, 0, sizeof(int), st));
    // (the column counters and the job boards are zeroed by the a1 / scan kernels: no per-fit memsets)

    // ---- a1 prologue ----
    const long long ntr = (long long)a->B * a->K;
    if (a->psc_dev) {
        if (a->T <= 0) { set_error("cm_caviar_fit: T=%d", a->T); return CM_ESHAPE; }
        const unsigned blocks = (unsigned)((ntr + 7) / 8);
        if (a->psc_dtype == CM_F32) psc_stats_kernel<float><<<blocks, 256, 0, st>>>((const float*)a->psc_dev, ntr, a->T, L, ws, a->K);
        else if (a->psc_dtype == CM_F64) psc_stats_kernel<double><<<blocks, 256, 0, st>>>((const double*)a->psc_dev, ntr, a->T, L, ws, a->K);
        else { set_error("cm_caviar_fit: bad psc dtype"); return CM_EINVAL; }
    } else {
        copy_stats_kernel<<<(unsigned)((ntr + 255) / 256), 256, 0, st>>>(a->y_dev, a->ss_dev, L, ws, a->K, ntr);
    }
    count_launch();
    PowerTable pt{};
    pt.P = a->n_powers;
    for (int i = 0; i < a->n_powers; ++i) pt.v[i] = a->powers[i];
    int rc;
    if (a->stim_dtype == CM_F32) rc = run_csr<float>(a, L, ws, pt, st);
    else if (a->stim_dtype == CM_F64) rc = run_csr<double>(a, L, ws, pt, st);
    else if (a->stim_dtype == CM_U8) rc = run_csr<unsigned char>(a, L, ws, pt, st);
    else { set_error("cm_caviar_fit: bad stim dtype"); return CM_EINVAL; }
    if (rc) return rc;

    // ---- the persistent fit kernel ----
    FitParams p{};
    p.L = L; p.ws = ws; p.B = a->B; p.N = a->N; p.K = a->K; p.P = a->n_powers; p.nnz_cap = a->nnz_cap;
    for (int i = 0; i < a->n_powers; ++i) p.powers[i] = a->powers[i];
    p.mu0 = a->mu0_dev; p.beta0 = a->beta0_dev; p.phi0 = a->phi0_dev; p.phicov0 = a->phi_cov0_dev;
    bool same = true;
    for (int b = 1; b < a->B; ++b) same = same && a->shape0[b] == a->shape0[0] && a->rate0[b] == a->rate0[0];
    if (!same) { set_error("cm_caviar_fit: per-fit shape/rate priors must currently be identical within one call"); return CM_EUNSUPPORTED; }
    p.shape0 = a->shape0[0]; p.rate0 = a->rate0[0]; p.shape0_arr = nullptr; p.rate0_arr = nullptr;
    p.seeds = seeds_dev; p.opt = a->opt;
    p.mu_out = a->mu_dev; p.beta_out = a->beta_dev; p.shape_out = a->shape_dev; p.rate_out = a->rate_dev;
    p.phi_out = a->phi_dev; p.phicov_out = a->phi_cov_dev; p.z_out = a->z_dev;
    p.mu_hist = a->mu_hist_dev; p.beta_hist = a->beta_hist_dev; p.shape_hist = a->shape_hist_dev;
    p.rate_hist = a->rate_hist_dev; p.phi_hist = a->phi_hist_dev; p.phicov_hist = a->phi_cov_hist_dev;
    p.z_hist = a->z_hist_dev; p.lamhist = want_lamhist ? 1 : 0;
    p.status = a->status_dev;
    const int smem_bytes = small_cta ? v256.smem_bytes : fit512::FIT_SMEM_BYTES;
    p.smem_doubles = smem_bytes / 8;
    // helper CTAs for the panel GEMMs of single large fits: only when every CTA of the launch is co-resident
    int ct = 1, sm_count = 0;
    {
        int dev = 0;
        CM_CUDA_CHECK(cudaGetDevice(&dev));
        CM_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    if (!small_cta && a->N > 256) {
        const int sms = sm_count;
        int want = a->N >= 2048 ? 127 : (a->N >= 768 ? 63 : 15);                  // large systems: the trailing updates / inverse columns of the tile solve and
                                                             // the O(K) passes scale with the CTA count (C5: 47 -> 127 helpers, 628 -> 572 ms)
        if (const char* e = getenv("CM_CAVIAR_HELPERS")) want = atoi(e);
        want = want < 0 ? 0 : (want > 127 ? 127 : want);
        const int room = sms / a->B - 1;                      // helpers per fit that still leave every CTA resident
        if (want > room) want = room;
        if (want > 0) ct = want + 1;
    }
    p.ct = ct;
    main_kernel_begin(st);
#define CM_LAUNCH_FIT(NS, PT)                                                                                      \
    do {                                                                                                           \
        CM_CUDA_CHECK(cudaFuncSetAttribute(NS::caviar_fit_kernel<PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           smem_bytes));                                                           \
        if (ct > 1) {       /* helpers spin on the job board: they must all be resident together */                \
            int occ = 0;                                                                                           \
            CM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, NS::caviar_fit_kernel<PT>, NS::NT,   \
                                                                        (size_t)smem_bytes));                      \
            if ((long long)a->B * ct > (long long)occ * sm_count) { ct = 1; p.ct = 1; }                            \
        }                                                                                                          \
        if (ct > 1) {       /* cooperative launch guarantees co-residency (or fails loudly) */                     \
            void* kargs[] = {(void*)&p};                                                                           \
            CM_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)NS::caviar_fit_kernel<PT>, dim3(a->B * ct),           \
                                                      dim3(NS::NT), kargs, (size_t)smem_bytes, st));               \
        } else {                                                                                                   \
            int occ = 0;                                                                                           \
            CM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, NS::caviar_fit_kernel<PT>, NS::NT,   \
                                                                        (size_t)smem_bytes));                      \
            const long long wave = (long long)(occ > 0 ? occ : 1) * sm_count;                                      \
            p.queue = queue_dev;                                                                                   \
            NS::caviar_fit_kernel<PT><<<(unsigned)(a->B < wave ? a->B : wave), NS::NT, smem_bytes, st>>>(p);       \
        }                                                                                                          \
    } while (0)
    if (small_cta) {
        const int e = fit256_launch(p, a->n_powers, a->B, sm_count, queue_dev, st);
        if (e != 0) { set_error("fit256 launch failed: %s", cudaGetErrorString((cudaError_t)e)); return CM_ECUDA; }
    } else {
        if (a->n_powers <= 4) CM_LAUNCH_FIT(fit512, 4); else CM_LAUNCH_FIT(fit512, PMAX);
    }
#undef CM_LAUNCH_FIT
    main_kernel_end(st);
    count_launch();
    CM_CUDA_CHECK(cudaGetLastError());

    // ---- dense lam outputs ----
    if (a->lam_dev) {
        CM_CUDA_CHECK(cudaMemsetAsync(a->lam_dev, 0, (size_t)a->B * a->N * a->K * 8, st));
        densify_kernel<<<dim3(a->N, 1, a->B), 128, 0, st>>>(L, ws, a->N, a->K, a->lam_dev, 0, a->status_dev);
        count_launch();
    }
    if (a->lam_csr_val_dev || a->lam_csr_col_dev || a->lam_csr_ptr_dev) {
        // sparse posterior: the kernel's own CSR (rows = neurons, columns = trials that pass the lam_mask); three strided
        // device-to-device copies, no kernel.  Entries past lam_csr_ptr[N] are unspecified.
        if (!(a->lam_csr_val_dev && a->lam_csr_col_dev && a->lam_csr_ptr_dev)) {
            set_error("cm_caviar_fit: lam_csr_val_dev, lam_csr_col_dev and lam_csr_ptr_dev go together");
            return CM_EINVAL;
        }
        const size_t z = (size_t)a->nnz_cap;
        CM_CUDA_CHECK(cudaMemcpy2DAsync(a->lam_csr_val_dev, z * 8, ws + L.lam, L.stride, z * 8, a->B, cudaMemcpyDeviceToDevice, st));
        CM_CUDA_CHECK(cudaMemcpy2DAsync(a->lam_csr_col_dev, z * 4, ws + L.col_k, L.stride, z * 4, a->B, cudaMemcpyDeviceToDevice, st));
        CM_CUDA_CHECK(cudaMemcpy2DAsync(a->lam_csr_ptr_dev, (size_t)(a->N + 1) * 4, ws + L.row_ptr, L.stride,
                                        (size_t)(a->N + 1) * 4, a->B, cudaMemcpyDeviceToDevice, st));
    }
    if (want_lamhist) {
        CM_CUDA_CHECK(cudaMemsetAsync(a->lam_hist_dev, 0, (size_t)a->B * a->opt.iters * a->N * a->K * 8, st));
        densify_kernel<<<dim3(a->N, a->opt.iters, a->B), 128, 0, st>>>(L, ws, a->N, a->K, a->lam_hist_dev, a->opt.iters, a->status_dev);
        count_launch();
    }
    CM_CUDA_CHECK(cudaGetLastError());
    return CM_OK;
}
