// Compile-time plan of the multi-trace tensor-core demixer (csrc/nwd_mt.cu), shared by the device code and the
// host-side weight packer.  Network: circuitmap/neural_waveform_demixing.py:254-287 (layer shapes :259-269).
//
// Every convolution of the U-Net is ONE implicit GEMM shape on tcgen05 (kind::f16, fp16 operands, fp32 accumulate):
//
//     D[rho][(n, co)] = sum_v sum_ci  X[PH * rho + v][ci] * WS[v + n][ci][co]            (1)
//
// * a GEMM row `rho` is a group of PH consecutive output positions ("phases") of one trace; rows of G traces are
//   stacked so that one 128-row M tile carries several traces;
// * the N dimension is (phase n, output channel co), N = PH * COUT (16..128): the weight operand is the plain tap
//   table WS[u][co][8 ci] read at a sliding 16-byte offset (u = v + n), so the Toeplitz expansion costs no memory;
// * the output position of column block n is t = PH * q + (PH - 1 - n), WS[u] = w[u - (PH - 1)] (zero outside the taps);
// * activations are stored phase-split, XP[ci / 8][pos % PH][rho][8 ci] (16-byte units), which makes the A rows of
//   (1) consecutive 16-byte units = the no-swizzle K-major core-matrix layout of the UMMA shared-memory descriptor,
//   with a window offset v = PH * a + j being nothing but (phase plane j, unit offset a).
#pragma once
#include <cstdint>

namespace cm {
namespace nwdmt {

constexpr int G = 4;                 // traces per CTA pass
constexpr int THREADS = 512;
constexpr int T = 900;
constexpr int NLAYER = 9;            // d1 d2 d3 d4 u1 u2 u3 u4 fin

constexpr int L_P1 = 449, L_E1 = 387, L_P2 = 193, L_E2 = 162, L_P3 = 80, L_E3 = 65, L_P4 = 32, L_E4 = 17;
constexpr int L_U1 = 32, L_U2 = 80, L_U3 = 193, L_U4H = 402, L_U4 = 804;

struct LayerCfg {
    int PH;      // output positions per GEMM row
    int CIN;     // GEMM input channels per window position (8 = one 16-byte unit, K step = two window positions)
    int COUT;    // GEMM output channels per phase
    int TAPS;    // window length in positions of the (possibly re-grouped) input sequence
    int Q;       // GEMM rows per sequence (= units per phase plane per sequence)
    int SEQ;     // sequences per trace (2 where the input is split by parity: d1, fin)
    int UPAD;    // tap-table length U (>= TAPS + 2 PH - 2)
};
// d1: 1 input channel, dilation 2 -> two parity sequences of the pooled input; a "position" is a group of 8 samples
//     and the 128 GEMM output channels are (sub-position m, co) = 16 m + co.
// u4: stride-2 transposed convolution -> output channels (parity, co) = 4 par + co over input positions.
// fin: k = 256, dilation 2, 4 -> 1 channels: two outer parity sequences; a "position" is a pair of samples of one
//     parity sequence x 4 channels, output channels = (inner parity of the output index).
__host__ __device__ constexpr LayerCfg lcfg(int l) {
    constexpr LayerCfg tab[NLAYER] = {
        {1, 8, 128, 6, 29, 2, 6},      // d1
        {8, 16, 16, 32, 25, 1, 46},    // d2
        {4, 16, 32, 16, 20, 1, 22},    // d3
        {1, 32, 32, 16, 32, 1, 16},    // d4
        {2, 32, 16, 16, 24, 1, 18},    // u1
        {4, 48, 16, 16, 24, 1, 22},    // u2
        {8, 32, 16, 32, 28, 1, 46},    // u3
        {16, 32, 8, 16, 27, 1, 46},    // u4
        {32, 8, 2, 129, 12, 2, 192},   // fin
    };
    return tab[l];
}
__host__ __device__ constexpr int lc_v(int l) { return lcfg(l).TAPS + lcfg(l).PH - 1; }          // window positions
__host__ __device__ constexpr int lc_n(int l) { return lcfg(l).PH * lcfg(l).COUT; }
__host__ __device__ constexpr int lc_rows(int l) { return G * lcfg(l).SEQ * lcfg(l).Q; }
__host__ __device__ constexpr int lc_tiles(int l) { return (lc_rows(l) + 127) / 128; }
__host__ __device__ constexpr int lc_wbytes(int l) { return (lcfg(l).CIN / 8) * lcfg(l).UPAD * lcfg(l).COUT * 16; }
__host__ __device__ constexpr int lc_woff(int l) { int o = 0; for (int i = 0; i < l; ++i) o += lc_wbytes(i); return o; }
__host__ __device__ constexpr int lc_amax(int l) { return (lc_v(l) - 1) / lcfg(l).PH; }          // largest unit offset a
constexpr int W_BYTES = lc_woff(NLAYER);
constexpr int BIAS_OFF = W_BYTES;                       // 9 x 32 fp32 biases behind the fp16 weights
constexpr int BLOB_BYTES = BIAS_OFF + NLAYER * 32 * 4;

}  // namespace nwdmt
}  // namespace cm
