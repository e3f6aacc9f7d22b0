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
#include "common.cuh"
#include <vector>
#include <cfloat>
#include <math_constants.h>
#include <cstring>
#include <cstdlib>

namespace cm {
namespace cav {

constexpr int PMAX = CM_CAVIAR_MAX_POWERS;
constexpr int NB = 32;           // block size of the bordered Cholesky/inverse
constexpr int RG = 4;            // row groups of 8 in the panel GEMMs
constexpr int MAX_SHUFFLE_ROUNDS = 4;
constexpr int GK = 16, NST = 4;                                   // panel GEMM: k-chunk, pipeline stages
constexpr int XD_LD = NB + 4;
constexpr int ROWPAD = 2 + GK;                                    // slack rows so that whole chunks can be copied
constexpr int KSEG_MAX = 4;                                       // k segments of a panel GEMM (split-k over the CTAs of a fit)

// The 32-row panels (A block row, Lrow, W) are stored transposed, [column][32 rows], same bit-3 swizzle on odd columns.
__device__ __forceinline__ size_t pidx(int r, int cc) { return (size_t)cc * NB + (r ^ ((cc & 1) << 3)); }


// ------------------------------------------------------------------------------------------------ layout
struct Layout {
    size_t stride;
    // fp64
    size_t X, XI, Dinv, PA, PB, PP, lam, cst, y, ss, pred, resid, z, mu, beta, bvec, dvec, wvec, slam, slam2, sp, phibar, phi,
        phicov, phiz, phicovz, lamhist, lamT, growbuf, cscq, rcnt, mce;
    // int32 / uint32
    size_t row_ptr, col_ptr, colfill, col_k, csc_row, csc_pos, cntp, n0p, n1p, act, ainv, order, order2, pos, rownz,
        phizok, sortkeys, keys, dcnt, dlist, colpw, nmask, chinfo, ccol_ptr, ccsc_row, ccsc_pos, member, rowcb;
    // bytes
    size_t pw, mask, blocked;
    size_t job;      // job board of the panel-GEMM helper CTAs (ints)
};

static Layout make_layout(int N, int K, int64_t nnz, int iters, bool lamhist, int GCT, int NW) {
    Layout L{};
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~size_t(255); return r; };
    const size_t n = N, k = K, z = (size_t)nnz;
    const size_t npad = (n + 31) & ~size_t(31);
    {
        const size_t tiled = ((n + GCT - 1) / GCT) * (size_t)GCT * (n + ROWPAD) * 8, square = npad * npad * 8;
        L.X = take(tiled > square ? tiled : square);             // tiled X of the bordered recursion / square A of the tile solve
    }
    L.XI = take(GCT == 256 ? npad * npad * 8 : 0);               // inverse factor of the tile solve (16-warp variant only)
    L.Dinv = take(GCT == 256 ? npad * 32 * 8 : 0);               // inverses of its diagonal tiles
    L.PA = take((size_t)NB * (n + ROWPAD) * 8);
    L.PB = take((size_t)NB * (n + ROWPAD) * 8);
    L.PP = take((size_t)KSEG_MAX * NB * (n + ROWPAD) * 8);      // partial panels of the split-k panel GEMMs
    L.growbuf = take((size_t)(NW > NB ? NW : NB) * (n + 2) * 8);     // one row buffer per row of a 32-row block
    L.cscq = take(z * 16);
    L.lam = take(z * 8);
    L.cst = take(z * 8);
    L.y = take(k * 8);
    L.ss = take(k * 8);
    L.pred = take(k * 8);
    L.resid = take(k * 8);
    L.z = take(k * 8);
    L.mu = take(n * 8);
    L.beta = take(n * 8);
    L.bvec = take(n * 8);
    L.dvec = take(n * 8);
    L.wvec = take(n * 8);
    L.slam = take(n * 8);
    L.slam2 = take(n * 8);
    L.sp = take(n * PMAX * 8);
    L.phibar = take(n * 2 * 8);
    L.phi = take(n * 2 * 8);
    L.phicov = take(n * 4 * 8);
    L.phiz = take(n * 2 * 8);
    L.phicovz = take(n * 4 * 8);
    L.lamhist = take(lamhist ? (size_t)iters * z * 8 : 0);
    L.lamT = take(z * 8);
    L.rcnt = take(n * PMAX * 8);
    L.mce = take(n * PMAX * 8);
    L.row_ptr = take((n + 1) * 4);
    L.col_ptr = take((k + 1) * 4);
    L.colfill = take(k * 4);
    L.col_k = take(z * 4);
    L.csc_row = take(z * 4);
    L.csc_pos = take(z * 4);
    L.cntp = take(n * PMAX * 4);
    L.n0p = take(n * PMAX * 4);
    L.n1p = take(n * PMAX * 4);
    L.act = take(n * 4);
    L.ainv = take(n * 4);
    L.order = take(n * 4);
    L.order2 = take(n * 4);
    L.pos = take(n * 4);
    L.rownz = take(n * 4);
    L.phizok = take(n * 4);
    L.sortkeys = take(n * 4);
    L.keys = take(2 * n * 2 * 4);
    L.dcnt = take(n * 4);
    L.dlist = take(n * 4);
    L.colpw = take(z * 4);
    L.nmask = take(n * PMAX * 4);
    L.chinfo = take(n * 16);
    L.ccol_ptr = take((k + 1) * 4);
    L.ccsc_row = take(z * 4);
    L.ccsc_pos = take(z * 4);
    L.member = take(n * 4);
    L.rowcb = take(z * 8);
    L.pw = take(z);
    L.mask = take(k);
    L.blocked = take(k);
    L.job = take(256);
    L.stride = o;
    return L;
}

struct FitParams {
    Layout L;
    char* ws;
    int B, N, K, P;
    int64_t nnz_cap;
    double powers[PMAX];
    const double *mu0, *beta0, *phi0, *phicov0;
    double shape0, rate0;             // per-launch scalars when all fits share them, else arrays below
    const double *shape0_arr, *rate0_arr;
    const unsigned long long* seeds;  // device, B
    cm_caviar_options opt;
    double *mu_out, *beta_out, *shape_out, *rate_out, *phi_out, *phicov_out, *z_out;
    double *mu_hist, *beta_hist, *shape_hist, *rate_hist, *phi_hist, *phicov_hist, *z_hist;
    int lamhist;
    int* status;
    int smem_doubles;                 // dynamic shared memory available for pred / row buffers
    int ct;                           // CTAs per fit: 1 + panel-GEMM helpers (single large fits only)
    int* queue;                       // device counter handing out fit indices (nullptr: CTA b / ct runs fit b)
};

// ------------------------------------------------------------------------------------------------ PRNG
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// Threefry-2x32-20 (Random123), the block function behind jax.random (oracle/prng.py).
__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
    const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
    x0 += k0; x1 += k1;
#define TF_R(r) x0 += x1; x1 = rotl32(x1, r); x1 ^= x0;
    TF_R(13) TF_R(15) TF_R(26) TF_R(6)
    x0 += k1; x1 += k2 + 1u;
    TF_R(17) TF_R(29) TF_R(16) TF_R(24)
    x0 += k2; x1 += k0 + 2u;
    TF_R(13) TF_R(15) TF_R(26) TF_R(6)
    x0 += k0; x1 += k1 + 3u;
    TF_R(17) TF_R(29) TF_R(16) TF_R(24)
    x0 += k1; x1 += k2 + 4u;
    TF_R(13) TF_R(15) TF_R(26) TF_R(6)
    x0 += k2; x1 += k0 + 5u;
#undef TF_R
}

// jax.random.split(key): rows (new_key, subkey).  Executed by lanes 0 and 1 of a warp; every lane gets the result.
__device__ __forceinline__ void warp_split(uint32_t k0, uint32_t k1, uint32_t& r00, uint32_t& r01, uint32_t& r10,
                                           uint32_t& r11) {
    const int lane = threadIdx.x & 31;
    uint32_t x0 = lane & 1, x1 = 2 + (lane & 1);
    threefry2x32(k0, k1, x0, x1);
    r00 = __shfl_sync(0xffffffffu, x0, 0);
    r01 = __shfl_sync(0xffffffffu, x0, 1);
    r10 = __shfl_sync(0xffffffffu, x1, 0);
    r11 = __shfl_sync(0xffffffffu, x1, 1);
}

__device__ __forceinline__ double bits_to_unit_double(uint32_t hi, uint32_t lo) {
    const unsigned long long b = ((unsigned long long)hi << 32) | lo;
    return __longlong_as_double((long long)((b >> 12) | 0x3FF0000000000000ull)) - 1.0;
}

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

// 1 / (1 + exp(-x)) for the sweep, whose sequential chain is bound by the dependent-issue latency of this expression
// (caviar.py:216): exp by argument reduction + a degree-13 polynomial in Estrin form (depth 4 instead of 13), the quotient
// by the hardware reciprocal seed + three Newton steps instead of the IEEE division sequence.  Error <= 1 ulp, the same
// class as the library call it replaces (CUDA's exp is not correctly rounded either); arguments beyond +-700 -- where exp
// overflows or underflows and the reference's result is exactly 0 or 1 -- take the library path, so those exact values
// (which update_phi's nan_to_num handling depends on) are produced by the same instructions as before.
__device__ __noinline__ double sigmoid_edge(double t) { return 1.0 / (1.0 + exp(t)); }   // out of line: keeps the sweep's loop body small
__device__ __forceinline__ double sigmoid_fast(double x) {
    const double t = -x;
    if (!(fabs(t) <= 700.0)) return sigmoid_edge(t);
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
    const double pe = fma(u1, r8, u0);
    const double e = pe * __longlong_as_double(((long long)((int)nf + 1023)) << 52);
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

// isotonic_regression(sr)[-1] with unit weights: mean of the last pool (pava.py:9-61).
__device__ __forceinline__ double pava_last(const double* sr, int P) {
    double v[PMAX], w[PMAX];
    int top = 0;
    v[0] = sr[0]; w[0] = 1.0;
    for (int t = 1; t < P; ++t) {
        ++top;
        v[top] = sr[t]; w[top] = 1.0;
        // x / 1.0 == x exactly, so singleton pools skip the division
        while (top > 0 && ((w[top - 1] == 1.0 ? v[top - 1] : v[top - 1] / w[top - 1]) >
                           (w[top] == 1.0 ? v[top] : v[top] / w[top]))) {
            --top;
            v[top] = v[top] + v[top + 1];
            w[top] = w[top] + w[top + 1];
        }
    }
    return w[top] == 1.0 ? v[top] : v[top] / w[top];
}

__device__ long long g_phase_cycles[32];
__device__ int g_phase_enable = 0;

// register-resident variant for the sweep's critical path: every array index is static after unrolling
template <int PT>
__device__ __forceinline__ double rget(const double (&a)[PT], int i) {
    double r = a[0];
#pragma unroll
    for (int p = 1; p < PT; ++p) r = (i == p) ? a[p] : r;
    return r;
}
template <int PT>
__device__ __forceinline__ void rset(double (&a)[PT], int i, double x) {
#pragma unroll
    for (int p = 0; p < PT; ++p) a[p] = (i == p) ? x : a[p];
}
template <int PT>
__device__ __forceinline__ double pava_last_reg(const double (&sr)[PT], int P) {
    double v[PT], w[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) { v[p] = 0.0; w[p] = 1.0; }
    int top = 0;
    v[0] = sr[0];
#pragma unroll
    for (int t = 1; t < PT; ++t) {
        if (t < P) {
            ++top;
            rset<PT>(v, top, sr[t]);
            rset<PT>(w, top, 1.0);
            while (top > 0) {
                const double vb = rget<PT>(v, top - 1), wb = rget<PT>(w, top - 1);
                const double vt = rget<PT>(v, top), wt = rget<PT>(w, top);
                const double mb = (wb == 1.0) ? vb : vb / wb;      // x / 1.0 == x exactly
                const double mt = (wt == 1.0) ? vt : vt / wt;
                if (!(mb > mt)) break;
                --top;
                rset<PT>(v, top, vb + vt);
                rset<PT>(w, top, wb + wt);
            }
        }
    }
    const double vt = rget<PT>(v, top), wt = rget<PT>(w, top);
    return (wt == 1.0) ? vt : vt / wt;
}

}  // namespace cav
}  // namespace cm

#define CM_NT 512
#define CM_FITNS fit512
namespace cm { namespace cav {
#include "caviar_fit.inl"
} }
#undef CM_NT
#undef CM_FITNS
#define CM_NT 256
#define CM_FITNS fit256
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

extern "C" size_t cm_caviar_workspace_bytes(int B, int N, int K, int64_t nnz_cap, int save_histories) {
    if (B <= 0 || N <= 0 || K <= 0 || nnz_cap < 0) return 0;
    // sized for the 16-warp layout (256-column tiles, 16 growbufs), which bounds the 8-warp one
    const Layout L = make_layout(N, K, nnz_cap, save_histories > 0 ? save_histories : 0, save_histories > 0, 256, 16);
    return L.stride * (size_t)B + (size_t)B * 8 + 512;      // + device copy of the seeds + fit queue counter
}

template <typename TS>
static int run_csr(const cm_caviar_args* a, const Layout& L, char* ws, const PowerTable& pt, cudaStream_t st) {
    dim3 grid(a->N, a->B);
    csr_count_kernel<TS><<<grid, 256, 0, st>>>((const TS*)a->stim_dev, a->N, a->K, L, ws, pt, a->status_dev,
                                               a->opt.y_xcorr_thresh);
    scan_kernel<<<a->B, 1024, 0, st>>>(a->N, a->K, L, ws, (long long)a->nnz_cap, a->status_dev);
    csr_fill_kernel<TS><<<grid, 256, 0, st>>>((const TS*)a->stim_dev, a->N, a->K, L, ws, pt, a->status_dev,
                                              a->opt.y_xcorr_thresh);
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
    const Layout L = small_cta ? make_layout(a->N, a->K, a->nnz_cap, a->opt.iters, want_lamhist, fit256::GCT, fit256::NW)
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
    const int smem_bytes = small_cta ? fit256::FIT_SMEM_BYTES : fit512::FIT_SMEM_BYTES;
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
        int want = a->N >= 2048 ? 47 : 15;                   // large systems have more panel-GEMM jobs (column tiles x k segments)
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
        if (a->n_powers <= 4) CM_LAUNCH_FIT(fit256, 4); else CM_LAUNCH_FIT(fit256, PMAX);
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
