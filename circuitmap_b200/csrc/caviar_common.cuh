// Shared definitions of the CAVIaR translation units (csrc/caviar.cu: host API, prologue kernels, 16-warp fit kernels;
// csrc/caviar_fit256.cu: 8-warp fit kernels -- two units so that the two big kernel families compile in parallel).
#pragma once
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

static inline Layout make_layout(int N, int K, int64_t nnz, int iters, bool lamhist, int GCT, int NW) {
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
static __device__ __noinline__ double sigmoid_edge(double t) { return 1.0 / (1.0 + exp(t)); }   // out of line: keeps the sweep's loop body small
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

static __device__ long long g_phase_cycles[32];
static __device__ int g_phase_enable = 0;

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

// Cholesky factor of a 32 x 32 SPD block held in shared memory (Sd, lower triangle) and the inverse of that factor (Xd),
// by ONE warp with the matrix in REGISTERS: lane r owns row r (statically indexed after full unrolling).
//   factor, column j: the pivot comes by one shuffle, every lane scales its own entry with rsqrt(pivot), the finished
//     column goes through 32 doubles of shared memory (rows 0 / 1 of Xd, double-buffered, one __syncwarp per column)
//     and is read back as broadcasts for the rank-1 update of the remaining columns.  The update runs over the full
//     row: what it leaves above the diagonal never reaches an entry on or below it and is masked at the store.
//   inverse, column `lane` of X = L^-1 by column-oriented forward substitution: x[r] = -acc[r] / L[r][r], then
//     acc[r2] += L[r2][r] x[r] for r2 > r -- independent FMAs against broadcast reads of the factor; x[r] = 0 for
//     r < lane needs no masking.
// ~3 k instructions with short dependency chains; the earlier version (each column broadcast lane by lane with 64-bit
// shuffles, select-masked updates, row-oriented substitution with a chain of r/2 dependent FMAs per row) had ~7 k and
// sat on the critical path of every 32 rows of every solve.
static __device__ __forceinline__ double potrf_lds(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a) : "memory");
    return v;
}
static __device__ __noinline__ void potrf32_warp(double (*Sd)[NB + 1], double (*Xd)[XD_LD], int nb) {
    const int lane = threadIdx.x & 31;
    const uint32_t sd_a = (uint32_t)__cvta_generic_to_shared(&Sd[0][0]), xd_a = (uint32_t)__cvta_generic_to_shared(&Xd[0][0]);
    double a[32];
#pragma unroll
    for (int q = 0; q < 32; ++q)
        a[q] = (lane < nb) ? ((q <= lane) ? Sd[lane][q] : 0.0) : ((q == lane) ? 1.0 : 0.0);   // rows >= nb: identity padding
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const double ajj = __shfl_sync(0xffffffffu, a[j], j);
        const double rs = rsqrt(ajj);
        const double l = (lane == j) ? ajj * rs : a[j] * rs;     // lanes < j: above the diagonal, never used
        a[j] = l;
        double* cb = &Xd[j & 1][0];
        cb[lane] = l;
        __syncwarp();
#pragma unroll
        for (int q = j + 1; q < 32; ++q) a[q] = fma(-l, potrf_lds(xd_a + (uint32_t)(((j & 1) * XD_LD + q) * 8)), a[q]);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 32; ++q) Sd[lane][q] = (q <= lane) ? a[q] : 0.0;     // the factor (zeros above the diagonal)
    __syncwarp();
    const double ipiv = 1.0 / Sd[lane][lane];
    double acc[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) acc[q] = 0.0;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const double ir = __shfl_sync(0xffffffffu, ipiv, r);      // 1 / L[r][r]
        const double xr = (r == lane) ? ir : ((r > lane) ? -acc[r] * ir : 0.0);
        Xd[r][lane] = (lane < nb && r < nb) ? xr : 0.0;
#pragma unroll
        for (int r2 = r + 1; r2 < 32; ++r2) acc[r2] = fma(potrf_lds(sd_a + (uint32_t)((r2 * (NB + 1) + r) * 8)), xr, acc[r2]);
    }
    __syncwarp();
}

}  // namespace cav
}  // namespace cm


namespace cm {
namespace cav {
// the 8-warp variant lives in its own translation unit (csrc/caviar_fit256.cu)
struct VariantInfo { int NT, NW, GCT, smem_bytes; };
VariantInfo fit256_info();
// launches fit256::caviar_fit_kernel<4 | PMAX> on one resident wave with the fit queue; returns a cudaError_t as int
int fit256_launch(FitParams& p, int n_powers, int B, int sm_count, int* queue_dev, cudaStream_t st);
// diagnostics counters of that unit: adds them to out32[32] (may be NULL), clears them, sets the enable word
int fit256_debug(long long* out32, int enable);
}  // namespace cav
}  // namespace cm
