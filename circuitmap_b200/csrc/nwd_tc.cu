// NWD demixer forward on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32, accumulators in TMEM.
//
// Same network and call as csrc/nwd.cu (reference circuitmap/neural_waveform_demixing.py:204-287), but the seven
// convolutions with 16..48 input channels (7.3 of the 8.4 MMAC per trace) run as IMPLICIT GEMMs:
//     M = 128 output positions (TMEM lanes),  N = C_out (16 / 32),  K = taps x C_in, 8 per MMA.
// Activations live in shared memory as planes of 4 channels, [plane][position][4] fp32 = 16 bytes per position, which
// is exactly the no-swizzle K-major "core matrix" row of the UMMA shared-memory descriptor: 8 consecutive positions
// are one 128-byte core matrix (SBO = 128 B), the next 4 channels are the next plane (LBO = plane stride), and a
// convolution tap is nothing but a 16-byte-aligned shift of the descriptor start address -- no im2col copy exists.
// Transposed convolutions are valid convolutions over zero-padded planes with flipped weights; the stride-2 one
// computes both output parities in one MMA (N = 2 x 8).  Weights are pre-arranged on the host in UMMA blocks and
// streamed per layer into shared memory with one cp.async.bulk each (overlapping the previous layer's epilogue).
// Epilogues read TMEM with tcgen05.ld (lane = position), add the folded BatchNorm bias, apply ReLU and write the
// next layer's planes; pooling / linear interpolation / concat stay on CUDA cores as float4 passes.
// The first (1 input channel) and last (1 output channel) convolutions keep the fp32 CUDA-core code of nwd.cu.
#include "nwd_common.cuh"
#include <cmath>
#include <cstring>

namespace cm {
namespace nwdtc {

constexpr int T = CM_NWD_T;
constexpr int L_P1 = 449, L_E1 = 387, L_P2 = 193, L_E2 = 162, L_P3 = 80, L_E3 = 65, L_P4 = 32, L_E4 = 17;
constexpr int L_U1 = 32, L_U2 = 80, L_U3 = 193, L_U4 = 804, L_U4H = 402;
constexpr int THREADS = 512;

// plane-layout buffers: [planes][positions][4 channels]
constexpr int D3_PL = 8, D3_TP = 417, D3_PAD = 15;     // dec3 = [up3 (planes 0-3) | enc1 (4-7)], zero pad 15 for u4
constexpr int D2_PL = 8, D2_TP = 224, D2_PAD = 31;     // dec2 = [up2 | enc2], zero pad 31 for u3
constexpr int D1_PL = 12, D1_TP = 95, D1_PAD = 15;     // dec1 = [up1 (0-3) | enc3 (4-11)], zero pad 15 for u2
constexpr int E4_PL = 8, E4_TP = 47, E4_PAD = 15;      // enc4, zero pad 15 for u1
// dec4 (4 channels x 900, zero padded by 255 on both sides: u = t + 255) feeds the final convolution (k=256, dil=2).
// Its even / odd subsequences xs_p[v] = xp[2 v + p] are stored split in 8 phases, XS[p][v % 8][v / 8][4 ch], so that
// the final convolution becomes two implicit GEMMs with 8 consecutive outputs of one parity per accumulator row.
constexpr int D4_PAD = 255;
constexpr int XS_TP = 92;                              // positions per phase plane (705 / 8 = 89 used)
constexpr int XS_FLOATS = 2 * 8 * XS_TP * 4;           // 5888 floats
constexpr int G_ROWS = 272;                            // Toeplitz tap table G[r][c] = W[c][r - 7], zero outside
constexpr int G_FLOATS = G_ROWS * 4;

constexpr int OFF_D3 = 0;
constexpr int OFF_D2 = OFF_D3 + D3_PL * D3_TP * 4;
constexpr int OFF_D1 = OFF_D2 + D2_PL * D2_TP * 4;
constexpr int OFF_E4 = OFF_D1 + D1_PL * D1_TP * 4;
constexpr int OFF_D4 = OFF_E4 + E4_PL * E4_TP * 4;
constexpr int OFF_G = OFF_D4 + XS_FLOATS;
constexpr int OFF_S = OFF_G + G_FLOATS;
constexpr int S_FLOATS = 3216;
constexpr int OFF_W = OFF_S + S_FLOATS;                // weight ring (one layer at a time)
constexpr int W_FLOATS = 16384;                        // 64 KB: the largest layers (d4, u3)
constexpr int OFF_OROW = OFF_W + 8192;                 // fp64 output row, behind the 32 KB of prefetched d2 weights
constexpr int SMEM_FLOATS = OFF_W + W_FLOATS;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + 128;
static_assert(OFF_W % 4 == 0 && OFF_S % 4 == 0 && OFF_D2 % 4 == 0, "16-byte alignment of operand buffers");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// tensor-core layers: d2 d3 d4 u1 u2 u3 u4
constexpr int NL = 7;
__host__ __device__ constexpr int tc_ci(int l) { return l == 0 ? 16 : l == 1 ? 16 : l == 2 ? 32 : l == 3 ? 32 : l == 4 ? 48 : 32; }
__host__ __device__ constexpr int tc_n(int l) { return (l == 1 || l == 2) ? 32 : 16; }   // MMA N (u4: 2 parities x 8, 4 useful each)
__host__ __device__ constexpr int tc_taps(int l) { return (l == 0 || l == 5) ? 32 : 16; }
__host__ __device__ constexpr int tc_ksteps(int l) { return tc_taps(l) * tc_ci(l) / 8; }
__host__ __device__ constexpr int tc_wfloats(int l) { return tc_ksteps(l) * tc_n(l) * 8; }
__host__ __device__ constexpr int tc_woff(int l) { int o = 0; for (int i = 0; i < l; ++i) o += tc_wfloats(i); return o; }
constexpr int TC_BIAS_OFF = tc_woff(NL);                       // 7 x 32 biases behind the weights
constexpr int TC_TOTAL = TC_BIAS_OFF + NL * 32;

// fp32 packed weights of nwd.cu used for the first and last convolution
constexpr int FW0 = 0, FB0 = 512, FW8 = 74388, FB8 = 75412;
constexpr int FIN_KSTEPS = 132;
constexpr int ACC_HALF = 64;                          // TMEM column offset of the second issuer's accumulators                        // window positions v = 0, 2, ..., 262

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);       // version 1, no swizzle
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ float tf32r(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float4 tf32r4(float4 v) { return make_float4(tf32r(v.x), tf32r(v.y), tf32r(v.z), tf32r(v.w)); }

__host__ __device__ constexpr uint32_t idesc_tf32(int n) {        // D = F32, A = B = TF32, both K-major, M = 128
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ long long g_nwd_cycles[16];
__device__ int g_nwd_prof = 0;
#define NWD_MARK(id)                                                                     \
    do {                                                                                 \
        if (g_nwd_prof && blockIdx.x == 0 && threadIdx.x == 0) {                         \
            const long long t_ = clock64();                                              \
            g_nwd_cycles[id] += t_ - tmark;                                              \
            tmark = t_;                                                                  \
        }                                                                                \
    } while (0)

// ---------------------------------------------------------------------------------------------- layer pieces
struct Pipe {
    uint64_t* bar_w;
    uint64_t* bar_mma;
    uint32_t wcount, mcount;        // completed phases waited so far (uniform across threads)
    uint32_t tmem;
};

// Issue all MMAs of one layer (one thread): A planes at `a_base` (floats, plane stride TP positions), weights in wbuf.
// Descriptors are built once per tile and advanced by constant deltas (the address field counts 16-byte units), so the
// issuing thread spends a handful of integer instructions per MMA.
// Two threads (warps 0 and 1) issue concurrently: each takes half of the taps and accumulates into its own TMEM
// columns (the epilogue adds the halves) -- one issuer alone is bound by its ~56-cycle issue latency, two reach the
// shared-memory operand bandwidth limit (~39 cycles per 128x16x8 MMA, measured in tests/tools/micro/tc_rate2.cu).
template <int L, int TP, int DIL, int NTILES>
__device__ __forceinline__ void issue_layer(const float* a_base, const float* wbuf, uint32_t tmem, int half) {
    constexpr int CI = tc_ci(L), N = tc_n(L), TAPS = tc_taps(L);
    constexpr int QP = CI / 8;                         // channel-group pairs per tap
    constexpr uint32_t idesc = idesc_tf32(N);
    const uint64_t a_desc0 = umma_desc(smem_u32(a_base), TP * 16, 128);
    const uint64_t b_desc0 = umma_desc(smem_u32(wbuf), (N / 8) * 128, 128);
#pragma unroll 1
    for (int mt = 0; mt < NTILES; ++mt) {
        const int j0 = half * (TAPS / 2);
        uint64_t ad = a_desc0 + (uint64_t)(128 * mt + j0 * DIL);   // +128 positions per tile, DIL per tap (16 B each)
        uint64_t bd = b_desc0 + (uint64_t)(j0 * QP * N * 2);
        const uint32_t td = tmem + half * ACC_HALF + mt * N;
#pragma unroll 4
        for (int j = 0; j < TAPS / 2; ++j) {
#pragma unroll
            for (int q = 0; q < QP; ++q) {
                umma_tf32(td, ad + (uint64_t)(2 * q * TP), bd + (uint64_t)(q * N * 2), idesc, (j | q) ? 1u : 0u);
            }
            ad += DIL;                                           // next tap: shift by DIL positions
            bd += QP * N * 2;                                    // next tap's weight blocks (N*32 bytes each)
        }
    }
}

// start the bulk copy of layer L's weight blocks into the shared weight buffer
__device__ __forceinline__ void load_weights(int l_off, int l_floats, const float* wtc, float* wbuf, uint64_t* bar) {
    mbar_expect_tx(bar, (uint32_t)l_floats * 4);
    bulk_g2s(wbuf, wtc + l_off, (uint32_t)l_floats * 4, bar);
}

// avg-pool (k=3, s=2) over planes, float4 per position
__device__ __forceinline__ void pool_planes(const float* in, int in_tp, int in_pad, float* out, int planes, int Lout) {
    for (int idx = threadIdx.x; idx < planes * Lout; idx += THREADS) {
        const int pl = idx / Lout, t = idx - pl * Lout;
        const float4* p = reinterpret_cast<const float4*>(in) + (size_t)pl * in_tp + in_pad + 2 * t;
        const float4 a = p[0], b = p[1], c = p[2];
        float4 r;
        r.x = (a.x + b.x + c.x) / 3.0f; r.y = (a.y + b.y + c.y) / 3.0f;
        r.z = (a.z + b.z + c.z) / 3.0f; r.w = (a.w + b.w + c.w) / 3.0f;
        reinterpret_cast<float4*>(out)[(size_t)pl * Lout + t] = tf32r4(r);
    }
}

// F.interpolate(linear, align_corners=False) over 4 planes (16 channels): S planes [4][Lin] -> dst planes 0..3
__device__ __forceinline__ void interp_planes(const float* in, int Lin, float* dst, int dst_tp, int dst_pad, int Lout) {
    const float scale = (float)Lin / (float)Lout;
    for (int idx = threadIdx.x; idx < 4 * Lout; idx += THREADS) {
        const int pl = idx / Lout, t = idx - pl * Lout;
        float src = scale * ((float)t + 0.5f) - 0.5f;
        src = src < 0.f ? 0.f : src;
        int i0 = (int)src;
        i0 = i0 < Lin - 1 ? i0 : Lin - 1;
        const int i1 = i0 + (i0 < Lin - 1 ? 1 : 0);
        const float l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f), l0 = 1.f - l1;
        const float4 a = reinterpret_cast<const float4*>(in)[(size_t)pl * Lin + i0];
        const float4 b = reinterpret_cast<const float4*>(in)[(size_t)pl * Lin + i1];
        float4 r;
        r.x = l0 * a.x + l1 * b.x; r.y = l0 * a.y + l1 * b.y; r.z = l0 * a.z + l1 * b.z; r.w = l0 * a.w + l1 * b.w;
        reinterpret_cast<float4*>(dst)[(size_t)pl * dst_tp + dst_pad + t] = tf32r4(r);
    }
}

// epilogue of a layer with N output channels in planes: TMEM tile rows = positions; bias + ReLU (+ tf32 rounding)
template <int N>
__device__ __forceinline__ void epilogue_planes(uint32_t tmem, int ntiles, const float* bias, float* dst, int dst_tp,
                                                int dst_pad, int plane0, int Lout) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = wid & 3;
    for (int mt = wid >> 2; mt < ntiles; mt += THREADS / 128) {
        const int t = 128 * mt + 32 * q + lane;
#pragma unroll
        for (int h = 0; h < N / 16; ++h) {
            uint32_t v[16], v2[16];
            tmem_ld16(tmem + mt * N + h * 16 + ((uint32_t)(32 * q) << 16), v);
            tmem_ld16(tmem + ACC_HALF + mt * N + h * 16 + ((uint32_t)(32 * q) << 16), v2);
            tmem_ld_wait();
            if (t < Lout) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float4 r;
                    r.x = fmaxf(__uint_as_float(v[4 * g + 0]) + __uint_as_float(v2[4 * g + 0]) + bias[h * 16 + 4 * g + 0], 0.f);
                    r.y = fmaxf(__uint_as_float(v[4 * g + 1]) + __uint_as_float(v2[4 * g + 1]) + bias[h * 16 + 4 * g + 1], 0.f);
                    r.z = fmaxf(__uint_as_float(v[4 * g + 2]) + __uint_as_float(v2[4 * g + 2]) + bias[h * 16 + 4 * g + 2], 0.f);
                    r.w = fmaxf(__uint_as_float(v[4 * g + 3]) + __uint_as_float(v2[4 * g + 3]) + bias[h * 16 + 4 * g + 3], 0.f);
                    reinterpret_cast<float4*>(dst)[(size_t)(plane0 + h * 4 + g) * dst_tp + dst_pad + t] = tf32r4(r);
                }
            }
        }
    }
}

__device__ __forceinline__ double block_max(double v, double* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = red[0];
    for (int i = 1; i < THREADS / 32; ++i) r = fmax(r, red[i]);
    return r;
}
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < THREADS / 32; ++i) r += red[i];
    return r;
}

// one TC layer: make A visible to the async proxy, issue MMAs, wait, prefetch the next layer's weights
template <int L, int TP, int DIL, int NTILES>
__device__ __forceinline__ void run_layer(Pipe& pp, const float* a_base, const float* wbuf, const float* wtc, int next_l) {
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0 || threadIdx.x == 32) {
        mbar_wait(pp.bar_w, pp.wcount & 1);
        tc_fence_after();
        issue_layer<L, TP, DIL, NTILES>(a_base, wbuf, pp.tmem, threadIdx.x >> 5);
        umma_commit(pp.bar_mma);
    }
    pp.wcount++;
    mbar_wait(pp.bar_mma, pp.mcount & 1);
    pp.mcount++;
    tc_fence_after();
    if (threadIdx.x == 0 && next_l >= 0) {
        int off = 0, fl = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i)
            if (i == next_l) { off = tc_woff(i); fl = tc_wfloats(i); }
        load_weights(off, fl, wtc, const_cast<float*>(wbuf), pp.bar_w);
    }
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(THREADS, 1)
nwd_forward_tc_kernel(const float* __restrict__ W, const float* __restrict__ Wtc, const TIn* __restrict__ traces,
                      TOut* __restrict__ outp, int K, int monotone_start, double* __restrict__ y_out,
                      double* __restrict__ ss_out) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ double red[THREADS / 32];
    __shared__ float bias_s[NL][32];
    float* d3 = smem + OFF_D3;
    float* d2 = smem + OFF_D2;
    float* d1 = smem + OFF_D1;
    float* e4 = smem + OFF_E4;
    float* d4 = smem + OFF_D4;
    float* S = smem + OFF_S;
    float* wbuf = smem + OFF_W;
    double* orow = reinterpret_cast<double*>(smem + OFF_OROW);
    const int wid = threadIdx.x >> 5;

    for (int i = threadIdx.x; i < OFF_S; i += THREADS) smem[i] = 0.f;       // zero pads are never written afterwards
    __syncthreads();
    for (int i = threadIdx.x; i < 256 * 4; i += THREADS) {                  // G[r][c] = W_final[c][r - 7] (tf32)
        const int j = i >> 2, c = i & 3;
        smem[OFF_G + (j + 7) * 4 + c] = tf32r(W[FW8 + c * 256 + j]);
    }
    for (int i = threadIdx.x; i < NL * 32; i += THREADS) bias_s[i / 32][i % 32] = Wtc[TC_BIAS_OFF + i];
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 2);                 // two MMA issuers commit per layer
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    Pipe pp;
    pp.bar_w = &bars[0]; pp.bar_mma = &bars[1]; pp.wcount = 0; pp.mcount = 0; pp.tmem = tmem_base_s;
    if (threadIdx.x == 0 && blockIdx.x < K) load_weights(tc_woff(0), tc_wfloats(0), Wtc, wbuf, pp.bar_w);

    long long tmark = clock64();
    for (int k = blockIdx.x; k < K; k += gridDim.x) {
        const bool more = k + (int)gridDim.x < K;
        const TIn* tr = traces + (size_t)k * T;
        // ---- normalise by the per-trace maximum (nwd.py:43-45), pool, first convolution in fp32 ----
        TIn v0 = tr[threadIdx.x];
        TIn v1 = (threadIdx.x + THREADS < T) ? tr[threadIdx.x + THREADS] : v0;
        const double tmax = block_max(fmax((double)v0, (double)v1), red);
        float* X = S;
        float* P1 = S + 900;
        X[threadIdx.x] = (float)(v0 / (TIn)tmax);
        if (threadIdx.x + THREADS < T) X[threadIdx.x + THREADS] = (float)(v1 / (TIn)tmax);
        __syncthreads();
        for (int t = threadIdx.x; t < L_P1; t += THREADS) P1[t] = (X[2 * t] + X[2 * t + 1] + X[2 * t + 2]) / 3.0f;
        __syncthreads();
        // d1: 1 -> 16 channels, 32 taps, dilation 2; thread = (plane, position), 4 channels each; enc1 -> dec3 planes 4..7
        for (int idx = threadIdx.x; idx < 4 * 416; idx += THREADS) {
            const int pl = idx / 416, t = idx - pl * 416;
            if (t >= L_E1) continue;
            float4 acc = *reinterpret_cast<const float4*>(W + FB0 + 4 * pl);
            const float* ip = P1 + t;
#pragma unroll 8
            for (int j = 0; j < 32; ++j) {
                const float x = ip[2 * j];
                const float4 wv = __ldg(reinterpret_cast<const float4*>(W + FW0 + j * 16 + 4 * pl));
                acc.x = fmaf(x, wv.x, acc.x); acc.y = fmaf(x, wv.y, acc.y);
                acc.z = fmaf(x, wv.z, acc.z); acc.w = fmaf(x, wv.w, acc.w);
            }
            acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
            reinterpret_cast<float4*>(d3)[(size_t)(4 + pl) * D3_TP + D3_PAD + t] = tf32r4(acc);
        }
        __syncthreads();
        NWD_MARK(0);
        // ---- encoder on tensor cores ----
        pool_planes(d3 + 4 * D3_TP * 4, D3_TP, D3_PAD, S, 4, L_P2);                       // pool2 -> S [4][193]
        run_layer<0, L_P2, 1, 2>(pp, S, wbuf, Wtc, 1);                                     // d2: 16->16, k32
        epilogue_planes<16>(pp.tmem, 2, bias_s[0], d2, D2_TP, D2_PAD, 4, L_E2);            // enc2 -> dec2 planes 4..7
        tc_fence_before();
        __syncthreads();
        NWD_MARK(1);
        pool_planes(d2 + 4 * D2_TP * 4, D2_TP, D2_PAD, S, 4, L_P3);                       // pool3 -> S [4][80]
        run_layer<1, L_P3, 1, 1>(pp, S, wbuf, Wtc, 2);                                     // d3: 16->32, k16
        epilogue_planes<32>(pp.tmem, 1, bias_s[1], d1, D1_TP, D1_PAD, 4, L_E3);            // enc3 -> dec1 planes 4..11
        tc_fence_before();
        __syncthreads();
        NWD_MARK(2);
        pool_planes(d1 + 4 * D1_TP * 4, D1_TP, D1_PAD, S, 8, L_P4);                       // pool4 -> S [8][32]
        run_layer<2, L_P4, 1, 1>(pp, S, wbuf, Wtc, 3);                                     // d4: 32->32, k16
        epilogue_planes<32>(pp.tmem, 1, bias_s[2], e4, E4_TP, E4_PAD, 0, L_E4);            // enc4 -> e4 planes 0..7
        tc_fence_before();
        __syncthreads();
        NWD_MARK(3);
        // ---- decoder: deconv (as padded valid conv) -> relu -> interp -> concat (up first) ----
        run_layer<3, E4_TP, 1, 1>(pp, e4, wbuf, Wtc, 4);                                   // u1: 32->16, k16
        epilogue_planes<16>(pp.tmem, 1, bias_s[3], S, L_U1, 0, 0, L_U1);
        tc_fence_before();
        __syncthreads();
        NWD_MARK(4);
        interp_planes(S, L_U1, d1, D1_TP, D1_PAD, L_E3);
        run_layer<4, D1_TP, 1, 1>(pp, d1, wbuf, Wtc, 5);                                   // u2: 48->16, k16
        epilogue_planes<16>(pp.tmem, 1, bias_s[4], S, L_U2, 0, 0, L_U2);
        tc_fence_before();
        __syncthreads();
        NWD_MARK(5);
        interp_planes(S, L_U2, d2, D2_TP, D2_PAD, L_E2);
        run_layer<5, D2_TP, 1, 2>(pp, d2, wbuf, Wtc, 6);                                   // u3: 32->16, k32
        epilogue_planes<16>(pp.tmem, 2, bias_s[5], S, L_U3, 0, 0, L_U3);
        tc_fence_before();
        __syncthreads();
        NWD_MARK(6);
        interp_planes(S, L_U3, d3, D3_TP, D3_PAD, L_E1);
        run_layer<6, D3_TP, 1, 4>(pp, d3, wbuf, Wtc, more ? 0 : -1);                       // u4: 32->4, k32, stride 2
        {
            // epilogue of u4: columns (parity p, co) = 8 p + co, rows = input index i; out position 2 i + p; scalar layout
            const int lane = threadIdx.x & 31, q = wid & 3;
            for (int mt = wid >> 2; mt < 4; mt += THREADS / 128) {
                uint32_t v[16], v2[16];
                tmem_ld16(pp.tmem + mt * 16 + ((uint32_t)(32 * q) << 16), v);
                tmem_ld16(pp.tmem + ACC_HALF + mt * 16 + ((uint32_t)(32 * q) << 16), v2);
                tmem_ld_wait();
                const int i = 128 * mt + 32 * q + lane;
                if (i < L_U4H) {
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int co = 0; co < 4; ++co)
                            S[co * L_U4 + 2 * i + p] =
                                fmaxf(__uint_as_float(v[8 * p + co]) + __uint_as_float(v2[8 * p + co]) + bias_s[6][co], 0.f);
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        NWD_MARK(7);
        {   // interp 804 -> 900 (nwd.py:237-238) straight into the parity / phase-split planes of the final convolution
            const float scale = (float)L_U4 / (float)T;
            for (int t = threadIdx.x; t < T; t += THREADS) {
                float src = scale * ((float)t + 0.5f) - 0.5f;
                src = src < 0.f ? 0.f : src;
                int i0 = (int)src;
                i0 = i0 < L_U4 - 1 ? i0 : L_U4 - 1;
                const int i1 = i0 + (i0 < L_U4 - 1 ? 1 : 0);
                const float l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f), l0 = 1.f - l1;
                float4 r;
                r.x = l0 * S[0 * L_U4 + i0] + l1 * S[0 * L_U4 + i1];
                r.y = l0 * S[1 * L_U4 + i0] + l1 * S[1 * L_U4 + i1];
                r.z = l0 * S[2 * L_U4 + i0] + l1 * S[2 * L_U4 + i1];
                r.w = l0 * S[3 * L_U4 + i0] + l1 * S[3 * L_U4 + i1];
                const int u = t + D4_PAD, par = u & 1, v = u >> 1;
                reinterpret_cast<float4*>(d4)[(size_t)(par * 8 + (v & 7)) * XS_TP + (v >> 3)] = tf32r4(r);
            }
        }
        NWD_MARK(8);
        // ---- final conv 4->1, k=256, dil=2, pad=255 (nwd.py:251-252, 285) as two Toeplitz GEMMs (one per output
        // parity): D[i][n] = sum_v sum_c xs_p[8 i + v][c] * G[v + n][c], output index tau = 8 i + 7 - n, t = 2 tau + p.
        proxy_fence();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (threadIdx.x == 0 || threadIdx.x == 32) {               // one issuer per output parity
            constexpr uint32_t idesc = idesc_tf32(16);
            const uint32_t g0 = smem_u32(smem + OFF_G);
            const int par = threadIdx.x >> 5;
            const uint64_t ad0 = umma_desc(smem_u32(d4 + (size_t)par * 8 * XS_TP * 4), XS_TP * 16, 128);
            uint64_t bd = umma_desc(g0, 16, 0);
            const uint32_t td = pp.tmem + par * 16;
#pragma unroll 1
            for (int s4 = 0; s4 < FIN_KSTEPS / 4; ++s4) {          // v = 8 s4 + {0, 2, 4, 6}: phases 0,2,4,6 at position s4
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    umma_tf32(td, ad0 + (uint64_t)(2 * e * XS_TP + s4), bd, idesc, (s4 | e) ? 1u : 0u);
                    bd += 2;                                     // G advances by two rows (32 bytes)
                }
            }
            umma_commit(pp.bar_mma);
        }
        mbar_wait(pp.bar_mma, pp.mcount & 1);
        pp.mcount++;
        tc_fence_after();
        {
            const int lane = threadIdx.x & 31, q = wid & 3;
            const float bf = __ldg(W + FB8);
            if (wid < 8) {                                     // warps 0..3: parity 0, warps 4..7: parity 1
                const int par = wid >> 2;
                uint32_t v[16];
                tmem_ld16(pp.tmem + par * 16 + ((uint32_t)(32 * q) << 16), v);
                tmem_ld_wait();
                const int i = 32 * q + lane;
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int t = 2 * (8 * i + 7 - n) + par;
                    if (t < T) {
                        const float o = fmaxf(__uint_as_float(v[n]) + bf, 0.f);
                        orow[t] = (double)((TOut)o * (TOut)tmax);
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        NWD_MARK(9);
        if (monotone_start >= 1 && monotone_start < T && threadIdx.x < 32) {
            double carry = orow[monotone_start - 1];
            for (int base = monotone_start; base < T; base += 32) {
                const int t = base + threadIdx.x;
                double v = t < T ? orow[t] : carry;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double u = __shfl_up_sync(0xffffffffu, v, o);
                    if ((int)threadIdx.x >= o) v = fmin(v, u);
                }
                v = fmin(v, carry);
                if (t < T) orow[t] = v;
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
        }
        __syncthreads();
        TOut* op = outp + (size_t)k * T;
        double s1 = 0.0, s2 = 0.0;
        for (int t = threadIdx.x; t < T; t += THREADS) {
            const double v = orow[t];
            op[t] = (TOut)v;
            s1 += v;
            s2 += v * v;
        }
        if (y_out != nullptr || ss_out != nullptr) {
            s1 = block_sum(s1, red);
            s2 = block_sum(s2, red);
            if (threadIdx.x == 0) {
                if (y_out) y_out[k] = s1 - 0.5 * (orow[0] + orow[T - 1]);
                if (ss_out) ss_out[k] = s2;
            }
        }
        __syncthreads();
        NWD_MARK(10);
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(pp.tmem), "r"(128u));
}

// ---------------------------------------------------------------------------------------------- host side
static float tf32_round_host(float x) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;     // round to nearest (ties away), 10-bit mantissa
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}

void pack_tc_weights(const float* const* t, std::vector<float>& out) {
    out.assign(TC_TOTAL, 0.f);
    // state_dict layer index of the tc layers: dblock2,3,4 = 1,2,3 ; ublock1..4 = 4,5,6,7
    const int sd_idx[NL] = {1, 2, 3, 4, 5, 6, 7};
    const int cout[NL] = {16, 32, 32, 16, 16, 16, 4};
    const int kk[NL] = {32, 16, 16, 16, 16, 32, 32};             // kernel size of the torch layer
    for (int l = 0; l < NL; ++l) {
        const float *w = t[6 * sd_idx[l]], *b = t[6 * sd_idx[l] + 1], *g = t[6 * sd_idx[l] + 2], *be = t[6 * sd_idx[l] + 3],
                    *rm = t[6 * sd_idx[l] + 4], *rv = t[6 * sd_idx[l] + 5];
        const int CI = tc_ci(l), N = tc_n(l), TAPS = tc_taps(l), QP = CI / 8, CO = cout[l], KW = kk[l];
        float* dst = out.data() + tc_woff(l);
        for (int co = 0; co < CO; ++co) {
            const double sc = (double)g[co] / std::sqrt((double)rv[co] + 1e-5);
            out[TC_BIAS_OFF + l * 32 + co] = (float)(((double)b[co] - (double)rm[co]) * sc + (double)be[co]);
        }
        for (int j = 0; j < TAPS; ++j)
            for (int q = 0; q < QP; ++q)
                for (int h = 0; h < 2; ++h)
                    for (int col = 0; col < N; ++col)
                        for (int c = 0; c < 4; ++c) {
                            const int ci = 8 * q + 4 * h + c;
                            double v = 0.0;
                            if (l <= 2) {                               // Conv1d weight (co, ci, k): B[(j,ci), co]
                                const double sc = (double)g[col] / std::sqrt((double)rv[col] + 1e-5);
                                v = (double)w[(col * CI + ci) * KW + j] * sc;
                            } else if (l <= 5) {                        // ConvTranspose1d (ci, co, k), flipped: tap j' = k-1-j
                                const double sc = (double)g[col] / std::sqrt((double)rv[col] + 1e-5);
                                v = (double)w[(ci * CO + col) * KW + (KW - 1 - j)] * sc;
                            } else {                                    // stride 2: column = 8 p + co, tap = p + 2 (15 - m')
                                const int p = col >> 3, co = col & 7;
                                if (co < CO) {
                                    const double sc = (double)g[co] / std::sqrt((double)rv[co] + 1e-5);
                                    v = (double)w[(ci * CO + co) * KW + (p + 2 * (15 - j))] * sc;
                                }
                            }
                            const size_t off = (size_t)(j * QP + q) * N * 8 + ((size_t)(h * (N / 8) + col / 8) * 8 + col % 8) * 4 + c;
                            dst[off] = tf32_round_host((float)v);
                        }
    }
}

}  // namespace nwdtc
}  // namespace cm
extern "C" int cm_nwd_debug_cycles(long long* out, int n, int enable) {
    using namespace cm::nwdtc;
    if (out && n > 0) {
        long long h[16];
        CM_CUDA_CHECK(cudaMemcpyFromSymbol(h, g_nwd_cycles, sizeof(h)));
        for (int i = 0; i < n && i < 16; ++i) out[i] = h[i];
    }
    long long z[16] = {0};
    CM_CUDA_CHECK(cudaMemcpyToSymbol(g_nwd_cycles, z, sizeof(z)));
    CM_CUDA_CHECK(cudaMemcpyToSymbol(g_nwd_prof, &enable, sizeof(int)));
    return CM_OK;
}
namespace cm {
namespace nwdtc {

template <typename TIn, typename TOut>
static int launch_t(cm_nwd* h, const void* in, void* out, int K, int ms, double* y, double* ss, cudaStream_t st) {
    auto kern = nwd_forward_tc_kernel<TIn, TOut>;
    CM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int grid = K < h->sm_count ? K : h->sm_count;
    main_kernel_begin(st);
    kern<<<grid, THREADS, SMEM_BYTES, st>>>(h->w_dev, h->wtc_dev, (const TIn*)in, (TOut*)out, K, ms, y, ss);
    main_kernel_end(st);
    count_launch();
    CM_CUDA_CHECK(cudaGetLastError());
    return CM_OK;
}

int launch(cm_nwd* h, const void* in, int in_dtype, void* out, int out_dtype, int K, int ms, double* y, double* ss,
           cudaStream_t st) {
    if (in_dtype == CM_F32 && out_dtype == CM_F32) return launch_t<float, float>(h, in, out, K, ms, y, ss, st);
    if (in_dtype == CM_F64 && out_dtype == CM_F64) return launch_t<double, double>(h, in, out, K, ms, y, ss, st);
    if (in_dtype == CM_F32 && out_dtype == CM_F64) return launch_t<float, double>(h, in, out, K, ms, y, ss, st);
    if (in_dtype == CM_F64 && out_dtype == CM_F32) return launch_t<double, float>(h, in, out, K, ms, y, ss, st);
    set_error("cm_nwd_forward: bad dtype %d/%d", in_dtype, out_dtype);
    return CM_EINVAL;
}

}  // namespace nwdtc
}  // namespace cm
