// NWD demixer forward for sm_100a -- whole U-Net fused per trace, activations resident in shared memory.
//
// Replaces NeuralDemixer.__call__ / NWDUNet.forward / _monotone_decay_filter
// (reference circuitmap/neural_waveform_demixing.py:36-54, 204-287, 337-348).
// This translation unit holds the fp32 CUDA-core path (bit-for-bit fp32 arithmetic, BN folded):
// one persistent CTA per SM walks over traces; every layer reads/writes shared memory only, HBM is
// touched once for the trace in and once for the demixed trace out (7.2 KB/trace fp32).
#include "nwd_common.cuh"
#include <vector>
#include <cmath>
#include <cstring>

namespace cm {
namespace nwd {

constexpr int T = CM_NWD_T;
// activation lengths
constexpr int L_P1 = 449, L_E1 = 387, L_P2 = 193, L_E2 = 162, L_P3 = 80, L_E3 = 65, L_P4 = 32, L_E4 = 17;
constexpr int L_U1 = 32, L_U2 = 80, L_U3 = 193, L_U4 = 804, L_U4H = 402;
// zero-padded concat buffers: transposed convs become valid convs over (k-1)-padded inputs
constexpr int D3_PAD = 15, D3_STRIDE = 417;   // dec3 = [up3(16) | enc1(16)] x 387, feeds u4 (16 taps per parity)
constexpr int D2_PAD = 31, D2_STRIDE = 224;   // dec2 = [up2(16) | enc2(16)] x 162, feeds u3 (k=32)
constexpr int D1_PAD = 15, D1_STRIDE = 95;    // dec1 = [up1(16) | enc3(32)] x 65,  feeds u2 (k=16)
constexpr int E4_PAD = 15, E4_STRIDE = 47;    // enc4 32 x 17, feeds u1 (k=16)
constexpr int D4_PAD = 255, D4_STRIDE = 1410; // dec4 4 x 900, feeds the final conv (k=256, dil=2, pad=255)

constexpr int OFF_D3 = 0;
constexpr int OFF_D2 = OFF_D3 + 32 * D3_STRIDE;
constexpr int OFF_D1 = OFF_D2 + 32 * D2_STRIDE;
constexpr int OFF_E4 = OFF_D1 + 48 * D1_STRIDE;
constexpr int OFF_D4 = OFF_E4 + 32 * E4_STRIDE;
constexpr int OFF_S = OFF_D4 + 4 * D4_STRIDE;
constexpr int S_FLOATS = 3216;                 // scratch: pooled inputs / deconv outputs (max 4 x 804)
constexpr int SMEM_FLOATS = OFF_S + S_FLOATS;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + T * 8 + 256;   // + fp64 output row + reduction scratch
constexpr int THREADS = 512;

// packed weight offsets (floats): Wt[ci][tap][co] then bias[co] (final conv bias padded to 4)
constexpr int WOFF_W[9] = {0, 528, 8736, 16960, 33376, 41584, 53888, 70288, 74388};
constexpr int WOFF_B[9] = {512, 8720, 16928, 33344, 41568, 53872, 70272, 74384, 75412};
constexpr int W0 = 0, B0 = 512;
constexpr int W1 = 528, B1 = 8720;
constexpr int W2 = 8736, B2 = 16928;
constexpr int W3 = 16960, B3 = 33344;
constexpr int W4 = 33376, B4 = 41568;
constexpr int W5 = 41584, B5 = 53872;
constexpr int W6 = 53888, B6 = 70272;
constexpr int W7 = 70288, B7 = 74384;
constexpr int W8 = 74388, B8 = 75412;
constexpr int W_TOTAL = 75416;

template <int CI, int CO, int KW, int DIL, int COT>
__device__ __forceinline__ void conv_relu(const float* __restrict__ in, int in_stride, const float* __restrict__ w,
                                          const float* __restrict__ bias, float* __restrict__ out, int out_stride,
                                          int Lout) {
    constexpr int NCT = CO / COT;
    const int Lp = (Lout + 31) & ~31;
    for (int idx = threadIdx.x; idx < NCT * Lp; idx += THREADS) {
        const int ct = idx / Lp;
        const int t = idx - ct * Lp;
        if (t >= Lout) continue;
        float acc[COT];
#pragma unroll
        for (int c = 0; c < COT; ++c) acc[c] = __ldg(bias + ct * COT + c);
        const float* wp = w + ct * COT;
        const float* ip = in + t;
        for (int ci = 0; ci < CI; ++ci) {
#pragma unroll 8
            for (int j = 0; j < KW; ++j) {
                const float x = ip[ci * in_stride + j * DIL];
                const float4* w4 = reinterpret_cast<const float4*>(wp + (ci * KW + j) * CO);
#pragma unroll
                for (int q = 0; q < COT / 4; ++q) {
                    const float4 wv = __ldg(w4 + q);
                    acc[4 * q + 0] = fmaf(x, wv.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(x, wv.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(x, wv.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(x, wv.w, acc[4 * q + 3]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < COT; ++c) out[(ct * COT + c) * out_stride + t] = fmaxf(acc[c], 0.f);
    }
}

// ConvTranspose1d(32->4, k=32, stride=2) as two 16-tap valid convs (one per output parity) over the padded input.
__device__ __forceinline__ void deconv_s2_relu(const float* __restrict__ in, const float* __restrict__ w,
                                               const float* __restrict__ bias, float* __restrict__ out) {
    constexpr int Lp = (L_U4H + 31) & ~31;
    for (int idx = threadIdx.x; idx < 2 * Lp; idx += THREADS) {
        const int par = idx / Lp;
        const int i = idx - par * Lp;
        if (i >= L_U4H) continue;
        float a0 = __ldg(bias + 0), a1 = __ldg(bias + 1), a2 = __ldg(bias + 2), a3 = __ldg(bias + 3);
        const float* ip = in + i;
        for (int ci = 0; ci < 32; ++ci) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const float x = ip[ci * D3_STRIDE + m];
                const float4 wv = __ldg(reinterpret_cast<const float4*>(w + ((ci * 16 + m) * 2 + par) * 4));
                a0 = fmaf(x, wv.x, a0); a1 = fmaf(x, wv.y, a1); a2 = fmaf(x, wv.z, a2); a3 = fmaf(x, wv.w, a3);
            }
        }
        const int t = 2 * i + par;
        out[0 * L_U4 + t] = fmaxf(a0, 0.f);
        out[1 * L_U4 + t] = fmaxf(a1, 0.f);
        out[2 * L_U4 + t] = fmaxf(a2, 0.f);
        out[3 * L_U4 + t] = fmaxf(a3, 0.f);
    }
}

template <int C>
__device__ __forceinline__ void avgpool(const float* __restrict__ in, int in_stride, float* __restrict__ out, int Lout) {
    for (int idx = threadIdx.x; idx < C * Lout; idx += THREADS) {
        const int c = idx / Lout, t = idx - c * Lout;
        const float* p = in + c * in_stride + 2 * t;
        out[c * Lout + t] = (p[0] + p[1] + p[2]) / 3.0f;
    }
}

// F.interpolate(mode='linear', align_corners=False) with PyTorch's fp32 index arithmetic.
template <int C>
__device__ __forceinline__ void interp(const float* __restrict__ in, int Lin, float* __restrict__ out, int out_stride,
                                       int Lout) {
    const float scale = (float)Lin / (float)Lout;
    for (int idx = threadIdx.x; idx < C * Lout; idx += THREADS) {
        const int c = idx / Lout, t = idx - c * Lout;
        float src = scale * ((float)t + 0.5f) - 0.5f;
        src = src < 0.f ? 0.f : src;
        int i0 = (int)src;
        i0 = i0 < Lin - 1 ? i0 : Lin - 1;
        const int i1 = i0 + (i0 < Lin - 1 ? 1 : 0);
        const float l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
        const float l0 = 1.f - l1;
        out[c * out_stride + t] = l0 * in[c * Lin + i0] + l1 * in[c * Lin + i1];
    }
}

__device__ __forceinline__ double block_reduce_max(double v, double* red) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double r = red[0];
    for (int i = 1; i < THREADS / 32; ++i) r = fmax(r, red[i]);
    return r;
}
__device__ __forceinline__ double block_reduce_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < THREADS / 32; ++i) r += red[i];
    return r;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(THREADS, 1)
nwd_forward_fp32_kernel(const float* __restrict__ W, const TIn* __restrict__ traces, TOut* __restrict__ outp, int K,
                        int monotone_start, double* __restrict__ y_out, double* __restrict__ ss_out) {
    extern __shared__ __align__(16) float smem[];
    float* d3 = smem + OFF_D3;
    float* d2 = smem + OFF_D2;
    float* d1 = smem + OFF_D1;
    float* e4 = smem + OFF_E4;
    float* d4 = smem + OFF_D4;
    float* S = smem + OFF_S;
    double* orow = reinterpret_cast<double*>(smem + SMEM_FLOATS);   // T doubles (8-byte aligned: SMEM_FLOATS even)
    double* red = orow + T;                                         // 32 doubles

    // zero everything once: the padding columns are never written afterwards
    for (int i = threadIdx.x; i < OFF_S; i += THREADS) smem[i] = 0.f;
    __syncthreads();

    for (int k = blockIdx.x; k < K; k += gridDim.x) {
        const TIn* tr = traces + (size_t)k * T;
        // ---- normalise by the per-trace maximum (nwd.py:43-45) ----
        TIn v0 = tr[threadIdx.x];
        TIn v1 = (threadIdx.x + THREADS < T) ? tr[threadIdx.x + THREADS] : v0;
        const double tmax = block_reduce_max(fmax((double)v0, (double)v1), red);
        float* X = S;                 // 900
        float* P1 = S + 900;          // 449
        X[threadIdx.x] = (float)(v0 / (TIn)tmax);
        if (threadIdx.x + THREADS < T) X[threadIdx.x + THREADS] = (float)(v1 / (TIn)tmax);
        __syncthreads();
        // ---- encoder (nwd.py:216-217, 273-276) ----
        avgpool<1>(X, T, P1, L_P1);
        __syncthreads();
        conv_relu<1, 16, 32, 2, 8>(P1, L_P1, W + W0, W + B0, d3 + 16 * D3_STRIDE + D3_PAD, D3_STRIDE, L_E1);
        __syncthreads();
        avgpool<16>(d3 + 16 * D3_STRIDE + D3_PAD, D3_STRIDE, S, L_P2);
        __syncthreads();
        conv_relu<16, 16, 32, 1, 4>(S, L_P2, W + W1, W + B1, d2 + 16 * D2_STRIDE + D2_PAD, D2_STRIDE, L_E2);
        __syncthreads();
        avgpool<16>(d2 + 16 * D2_STRIDE + D2_PAD, D2_STRIDE, S, L_P3);
        __syncthreads();
        conv_relu<16, 32, 16, 1, 4>(S, L_P3, W + W2, W + B2, d1 + 16 * D1_STRIDE + D1_PAD, D1_STRIDE, L_E3);
        __syncthreads();
        avgpool<32>(d1 + 16 * D1_STRIDE + D1_PAD, D1_STRIDE, S, L_P4);
        __syncthreads();
        conv_relu<32, 32, 16, 1, 4>(S, L_P4, W + W3, W + B3, e4 + E4_PAD, E4_STRIDE, L_E4);
        __syncthreads();
        // ---- decoder (nwd.py:231-238, 279-282): deconv -> relu -> interp -> concat (up first) ----
        conv_relu<32, 16, 16, 1, 4>(e4, E4_STRIDE, W + W4, W + B4, S, L_U1, L_U1);
        __syncthreads();
        interp<16>(S, L_U1, d1 + D1_PAD, D1_STRIDE, L_E3);
        __syncthreads();
        conv_relu<48, 16, 16, 1, 4>(d1, D1_STRIDE, W + W5, W + B5, S, L_U2, L_U2);
        __syncthreads();
        interp<16>(S, L_U2, d2 + D2_PAD, D2_STRIDE, L_E2);
        __syncthreads();
        conv_relu<32, 16, 32, 1, 4>(d2, D2_STRIDE, W + W6, W + B6, S, L_U3, L_U3);
        __syncthreads();
        interp<16>(S, L_U3, d3 + D3_PAD, D3_STRIDE, L_E1);
        __syncthreads();
        deconv_s2_relu(d3, W + W7, W + B7, S);
        __syncthreads();
        interp<4>(S, L_U4, d4 + D4_PAD, D4_STRIDE, T);
        __syncthreads();
        // ---- final conv 4->1, k=256, dil=2, pad=255 (nwd.py:251-252, 285) + rescale ----
        {
            const float* wf = W + W8;
            const float bf = __ldg(W + B8);
            for (int t = threadIdx.x; t < T; t += THREADS) {
                float a0 = bf, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                const float* ip = d4 + t;
#pragma unroll 8
                for (int j = 0; j < 256; ++j) {
                    a0 = fmaf(ip[0 * D4_STRIDE + 2 * j], __ldg(wf + 0 * 256 + j), a0);
                    a1 = fmaf(ip[1 * D4_STRIDE + 2 * j], __ldg(wf + 1 * 256 + j), a1);
                    a2 = fmaf(ip[2 * D4_STRIDE + 2 * j], __ldg(wf + 2 * 256 + j), a2);
                    a3 = fmaf(ip[3 * D4_STRIDE + 2 * j], __ldg(wf + 3 * 256 + j), a3);
                }
                const float o = fmaxf((a0 + a1) + (a2 + a3), 0.f);
                orow[t] = (double)((TOut)o * (TOut)tmax);     // nwd.py:46: f32 net output times tmax
            }
        }
        __syncthreads();
        // ---- monotone decay filter (nwd.py:337-343): running min seeded by column start-1 ----
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
            s1 = block_reduce_sum(s1, red);
            s2 = block_reduce_sum(s2, red);
            if (threadIdx.x == 0) {
                if (y_out) y_out[k] = s1 - 0.5 * (orow[0] + orow[T - 1]);     // unit-spacing trapezoid, caviar.py:28
                if (ss_out) ss_out[k] = s2;                                   // autocorrelation at lag 0, caviar.py:30
            }
        }
        __syncthreads();
    }
}

}  // namespace nwd
}  // namespace cm


using namespace cm;
using namespace cm::nwd;

// fold BN (eval) into conv weights in fp64 and pack as Wt[ci][tap][co]
static void pack_weights(const float* const* t, std::vector<float>& W) {
    W.assign(W_TOTAL, 0.f);
    struct L { int kind, ci, co, k; };   // kind 0 conv (co,ci,k), 1 deconv s1 (ci,co,k) flipped, 2 deconv s2, 3 final
    const L layers[9] = {{0, 1, 16, 32}, {0, 16, 16, 32}, {0, 16, 32, 16}, {0, 32, 32, 16}, {1, 32, 16, 16},
                         {1, 48, 16, 16}, {1, 32, 16, 32}, {2, 32, 4, 32}, {3, 4, 1, 256}};
    for (int l = 0; l < 9; ++l) {
        const float *w = t[6 * l], *b = t[6 * l + 1], *g = t[6 * l + 2], *be = t[6 * l + 3], *rm = t[6 * l + 4],
                    *rv = t[6 * l + 5];
        const L& s = layers[l];
        for (int co = 0; co < s.co; ++co) {
            const double sc = (double)g[co] / std::sqrt((double)rv[co] + 1e-5);
            W[WOFF_B[l] + co] = (float)(((double)b[co] - (double)rm[co]) * sc + (double)be[co]);
            for (int ci = 0; ci < s.ci; ++ci)
                for (int j = 0; j < s.k; ++j) {
                    if (s.kind == 0) {
                        W[WOFF_W[l] + (ci * s.k + j) * s.co + co] = (float)((double)w[(co * s.ci + ci) * s.k + j] * sc);
                    } else if (s.kind == 1) {      // out[t] = sum_j' xp[t+j'] * w[ci][co][k-1-j']
                        W[WOFF_W[l] + (ci * s.k + j) * s.co + co] =
                            (float)((double)w[(ci * s.co + co) * s.k + (s.k - 1 - j)] * sc);
                    } else if (s.kind == 2) {      // tap j = par + 2*(15-m')  ->  [ci][m'][par][co]
                        const int par = j & 1, m = 15 - (j >> 1);
                        W[WOFF_W[l] + ((ci * 16 + m) * 2 + par) * 4 + co] = (float)((double)w[(ci * s.co + co) * s.k + j] * sc);
                    } else {                       // final: [ci][j]
                        W[WOFF_W[l] + ci * 256 + j] = (float)((double)w[(co * s.ci + ci) * s.k + j] * sc);
                    }
                }
        }
    }
}

extern "C" int cm_nwd_create(const float* const* tensors, int n_tensors, cm_nwd_t** out) {
    if (!tensors || !out || n_tensors != CM_NWD_NUM_TENSORS) {
        set_error("cm_nwd_create: expected %d tensors, got %d", CM_NWD_NUM_TENSORS, n_tensors);
        return CM_EINVAL;
    }
    std::vector<float> W;
    pack_weights(tensors, W);
    cm_nwd* h = new cm_nwd();
    CM_CUDA_CHECK(cudaGetDevice(&h->device));
    CM_CUDA_CHECK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device));
    CM_CUDA_CHECK(cudaMalloc(&h->w_dev, W.size() * sizeof(float)));
    CM_CUDA_CHECK(cudaMemcpy(h->w_dev, W.data(), W.size() * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<float> Wtc;
    cm::nwdtc::pack_tc_weights(tensors, Wtc);
    CM_CUDA_CHECK(cudaMalloc(&h->wtc_dev, Wtc.size() * sizeof(float)));
    CM_CUDA_CHECK(cudaMemcpy(h->wtc_dev, Wtc.data(), Wtc.size() * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<unsigned char> Wmt;
    cm::nwdmt::pack_weights(tensors, Wmt);
    h->mt_ok = cm::nwdmt::weights_fit_fp16();
    CM_CUDA_CHECK(cudaMalloc(&h->wmt_dev, Wmt.size()));
    CM_CUDA_CHECK(cudaMemcpy(h->wmt_dev, Wmt.data(), Wmt.size(), cudaMemcpyHostToDevice));
    *out = h;
    return CM_OK;
}

extern "C" void cm_nwd_destroy(cm_nwd_t* h) {
    if (!h) return;
    cudaFree(h->w_dev);
    cudaFree(h->wtc_dev);
    cudaFree(h->wmt_dev);
    delete h;
}

template <typename TIn, typename TOut>
static int launch_fp32(cm_nwd_t* h, const void* in, void* out, int K, int ms, double* y, double* ss, cudaStream_t st) {
    auto kern = nwd_forward_fp32_kernel<TIn, TOut>;
    CM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int grid = K < h->sm_count ? K : h->sm_count;
    main_kernel_begin(st);
    kern<<<grid, THREADS, SMEM_BYTES, st>>>(h->w_dev, (const TIn*)in, (TOut*)out, K, ms, y, ss);
    main_kernel_end(st);
    count_launch();
    CM_CUDA_CHECK(cudaGetLastError());
    return CM_OK;
}

extern "C" int cm_nwd_forward(cm_nwd_t* h, const void* traces_dev, int in_dtype, void* out_dev, int out_dtype, int K,
                              int T_, int monotone_start, double* y_dev, double* ss_dev, void* stream) {
    reset_launch_count();
    if (T_ != CM_NWD_T) { set_error("cm_nwd_forward: T=%d unsupported (network is built for T=%d)", T_, CM_NWD_T); return CM_ESHAPE; }
    if (K < 0) { set_error("cm_nwd_forward: K=%d", K); return CM_ESHAPE; }
    if (K == 0) return CM_OK;
    if (!h || !traces_dev || !out_dev) { set_error("cm_nwd_forward: null argument"); return CM_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    if (h->precision == 2) return cm::nwdmt::launch(h, traces_dev, in_dtype, out_dev, out_dtype, K, monotone_start, y_dev, ss_dev, st);
    if (h->precision == 1) return cm::nwdtc::launch(h, traces_dev, in_dtype, out_dev, out_dtype, K, monotone_start, y_dev, ss_dev, st);
    if (in_dtype == CM_F32 && out_dtype == CM_F32) return launch_fp32<float, float>(h, traces_dev, out_dev, K, monotone_start, y_dev, ss_dev, st);
    if (in_dtype == CM_F64 && out_dtype == CM_F64) return launch_fp32<double, double>(h, traces_dev, out_dev, K, monotone_start, y_dev, ss_dev, st);
    if (in_dtype == CM_F32 && out_dtype == CM_F64) return launch_fp32<float, double>(h, traces_dev, out_dev, K, monotone_start, y_dev, ss_dev, st);
    if (in_dtype == CM_F64 && out_dtype == CM_F32) return launch_fp32<double, float>(h, traces_dev, out_dev, K, monotone_start, y_dev, ss_dev, st);
    set_error("cm_nwd_forward: bad dtype %d/%d", in_dtype, out_dtype);
    return CM_EINVAL;
}

extern "C" int cm_nwd_set_precision(cm_nwd_t* h, int precision) {
    if (!h || precision < 0 || precision > 2) { set_error("cm_nwd_set_precision: precision must be 0 (fp32), 1 (tf32) or 2 (fp16)"); return CM_EINVAL; }
    if (precision == 2 && !h->mt_ok) {
        set_error("cm_nwd_set_precision: the BatchNorm-folded weights of this network exceed the fp16 range; use mode 0 or 1");
        return CM_EINVAL;
    }
    h->precision = precision;
    return CM_OK;
}

extern "C" int cm_nwd_mt_debug_cycles(long long* out, int n, int enable) { return cm::nwdmt::debug_cycles(out, n, enable); }
extern "C" int cm_nwd_mt_debug_dump(void* dev_buf, int stage) { return cm::nwdmt::debug_dump(dev_buf, stage); }

/* test hook: the packed fp16 tap tables + biases of the multi-trace kernel (host memory, BLOB bytes) */
extern "C" int cm_nwd_mt_pack(const float* const* tensors, int n_tensors, unsigned char* out, size_t cap, size_t* need) {
    if (!tensors || n_tensors != CM_NWD_NUM_TENSORS) { set_error("cm_nwd_mt_pack: expected %d tensors", CM_NWD_NUM_TENSORS); return CM_EINVAL; }
    std::vector<unsigned char> W;
    cm::nwdmt::pack_weights(tensors, W);
    if (need) *need = W.size();
    if (out) {
        if (cap < W.size()) { set_error("cm_nwd_mt_pack: buffer too small"); return CM_EINVAL; }
        std::memcpy(out, W.data(), W.size());
    }
    return CM_OK;
}
