// Synthetic mapping experiments on the device (reference circuitmap/simulation.py:25-215, blockwise
// design, nreps = 1).  The reference draws from NumPy's global Mersenne-Twister stream WITHOUT seeding it, so there is no
// output to be identical to: this generator draws from the same DISTRIBUTIONS with a counter-based stream (threefry2x32
// keyed by the map's seed) and is tested on moments and design invariants (tests/test_simulate_gpu.py).
//
// Per map (B maps per call, one seed each):
//   * design (simulation.py:45-63): passes of a random neuron order cut into ceil(N/H) holograms, every hologram shown at
//     each power (highest first), until K trials exist; then the trial order is shuffled.  Both permutations are keyed
//     Feistel permutations (cycle-walking), so a trial's targets and a neuron's trials are closed-form -- no sort, no
//     K x N pass.
//   * neurons (:80-91,117-126): tau_r, tau_d, phi_0, phi_1 uniform; int(connection_prob N) connected cells, 20 % of them
//     strong (U[20,40]), the rest weak (Exp(4) + 9).
//   * spikes (:92-114): Bernoulli(sigmoid(phi_0 I - phi_1)) per stimulation; cells below 40 % spike rate at the top power
//     get random extra spikes up to it; latency 160 + Gamma(1e4 / I^2, 15); multiplicative noise LogNormal(0, 0.01).
//   * traces (:138-176): evoked PSC = bi-exponential kernel shifted to int(latency), unit area (+1e-5 guards as the
//     reference has them) x noise x weight; spontaneous PSCs in 5 % of the trials; correlated noise = a Gaussian process
//     with squared-exponential covariance (length 50, scale 4e-3), sampled as white noise filtered with the Gaussian
//     whose autocorrelation is that covariance (exact in the continuum; the reference samples the 900 x 900 covariance
//     with an SVD); iid noise sigma 6e-4.
// Kernels: one CTA per map (neurons), one warp per connected neuron (spikes), one CTA per trial (design column + trace).
// HBM-bound on the K x T trace write (4 or 8 bytes per sample).
#include "common.cuh"
#include <cmath>
#include <cstring>
#include <math_constants.h>

namespace cm {
namespace sim {

constexpr int PMAX = CM_CAVIAR_MAX_POWERS;
constexpr int FR = 192;                 // half-width of the GP filter (5.4 sigma at length 50)
constexpr int TMAX = 1024;              // samples per trace supported by the trace kernel
constexpr int TOPMAX = 1024;            // top-power stimulations per neuron supported by the spike kernel

struct Params {
    int N, K, T, H, P, nh, npass, n_conn, n_strong;
    double powers_desc[PMAX];           // descending: highest power first (simulation.py:48)
    unsigned char code_desc[PMAX];      // uint8 code of powers_desc[a] (index in the ascending table + 1)
    double min_latency, gamma_beta, sigma, strong_lo, strong_hi, weak_mean, min_weight, phi0_lo, phi0_hi, phi1_lo, phi1_hi,
        mult_noise_log_var, tau_r_min, tau_r_max, tau_delta_min, tau_delta_max, gp_scale, gp_lengthscale, spont_prob,
        max_power_min_spike_rate;
};

// per-map workspace layout (bytes)
struct Lay { size_t tau_r, tau_d, phi0, phi1, w, wrange, ev_time, ev_amp, stride; };
static Lay make_lay(int N, int K, int H) {
    Lay L{};
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += (b + 255) & ~size_t(255); return r; };
    L.tau_r = take((size_t)N * 8); L.tau_d = take((size_t)N * 8); L.phi0 = take((size_t)N * 8); L.phi1 = take((size_t)N * 8);
    L.w = take((size_t)N * 8); L.wrange = take(16);
    L.ev_time = take((size_t)K * H * 2); L.ev_amp = take((size_t)K * H * 4);
    L.stride = o;
    return L;
}

// ------------------------------------------------------------------------------------------------ counter-based RNG
__device__ __forceinline__ uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
__device__ __forceinline__ void tf2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
    const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
    x0 += k0; x1 += k1;
#define R_(r) x0 += x1; x1 = rotl(x1, r); x1 ^= x0;
    R_(13) R_(15) R_(26) R_(6)  x0 += k1; x1 += k2 + 1u;
    R_(17) R_(29) R_(16) R_(24) x0 += k2; x1 += k0 + 2u;
    R_(13) R_(15) R_(26) R_(6)  x0 += k0; x1 += k1 + 3u;
    R_(17) R_(29) R_(16) R_(24) x0 += k1; x1 += k2 + 4u;
    R_(13) R_(15) R_(26) R_(6)  x0 += k2; x1 += k0 + 5u;
#undef R_
}
// streams: independent purposes get independent keys
enum { S_TAU = 1, S_PHI, S_CONN, S_W, S_SPK, S_PAD, S_LAT, S_MULT, S_SPONT, S_GP, S_IID, S_PERMN = 64, S_PERMK = 63 };

struct Rng {
    uint32_t k0, k1;
    __device__ Rng(unsigned long long seed, uint32_t stream) : k0((uint32_t)seed ^ (stream * 0x9E3779B9u)), k1((uint32_t)(seed >> 32) + stream) {}
    __device__ __forceinline__ void bits(unsigned long long idx, uint32_t& a, uint32_t& b) const {
        a = (uint32_t)idx; b = (uint32_t)(idx >> 32);
        tf2x32(k0, k1, a, b);
    }
    __device__ __forceinline__ double u(unsigned long long idx) const {        // (0, 1)
        uint32_t a, b;
        bits(idx, a, b);
        return ((double)(((unsigned long long)a << 21) ^ (b >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
    }
    __device__ __forceinline__ float2 normal2f(unsigned long long idx) const {  // two standard normals (Box-Muller, fp32)
        uint32_t a, b;
        bits(idx, a, b);
        const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float r = sqrtf(-2.0f * __logf(u1));
        float s, c;
        __sincosf(6.283185307179586f * u2, &s, &c);
        return make_float2(r * c, r * s);
    }
    __device__ __forceinline__ double normal(unsigned long long idx) const {
        const double u1 = u(2 * idx), u2 = u(2 * idx + 1);
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
};

// keyed permutation of [0, M): 4-round Feistel network on 2 * hb bits with cycle walking; invertible
struct Perm {
    uint32_t k0, k1, M;
    int hb;
    __device__ Perm(unsigned long long seed, uint32_t stream, uint32_t M_) : k0((uint32_t)seed + 0x632BE5ABu * stream), k1((uint32_t)(seed >> 32) ^ stream), M(M_) {
        int b = 1;
        while ((1ull << (2 * b)) < (unsigned long long)M_) ++b;
        hb = b;
    }
    __device__ __forceinline__ uint32_t f(uint32_t r, uint32_t round) const {
        uint32_t a = r, b = round;
        tf2x32(k0, k1, a, b);
        return a & ((1u << hb) - 1u);
    }
    __device__ __forceinline__ uint32_t fwd1(uint32_t x) const {
        uint32_t l = x >> hb, r = x & ((1u << hb) - 1u);
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) { const uint32_t t = l ^ f(r, i); l = r; r = t; }
        return (l << hb) | r;
    }
    __device__ __forceinline__ uint32_t inv1(uint32_t x) const {
        uint32_t l = x >> hb, r = x & ((1u << hb) - 1u);
#pragma unroll
        for (int i = 3; i >= 0; --i) { const uint32_t t = r ^ f(l, (uint32_t)i); r = l; l = t; }
        return (l << hb) | r;
    }
    __device__ __forceinline__ uint32_t fwd(uint32_t x) const { do { x = fwd1(x); } while (x >= M); return x; }
    __device__ __forceinline__ uint32_t inv(uint32_t x) const { do { x = inv1(x); } while (x >= M); return x; }
};

// ------------------------------------------------------------------------------------------------ kernels
// neurons: kernel time constants, sigmoid coefficients, weights (simulation.py:80-91,117-126); one CTA per map
__global__ void __launch_bounds__(256) sim_neurons_kernel(const Params p, const Lay L, char* ws, const unsigned long long* seeds,
                                                          double* weights_out) {
    const int b = blockIdx.x;
    char* base = ws + (size_t)b * L.stride;
    const unsigned long long seed = seeds[b];
    double* tau_r = (double*)(base + L.tau_r); double* tau_d = (double*)(base + L.tau_d);
    double* phi0 = (double*)(base + L.phi0); double* phi1 = (double*)(base + L.phi1);
    double* w = (double*)(base + L.w);
    const Rng rt(seed, S_TAU), rp(seed, S_PHI), rw(seed, S_W);
    const Perm conn(seed, S_CONN, (uint32_t)p.N);
    double lo = CUDART_INF, hi = -CUDART_INF;
    for (int n = threadIdx.x; n < p.N; n += blockDim.x) {
        const double tr = p.tau_r_min + (p.tau_r_max - p.tau_r_min) * rt.u(2ull * n);
        tau_r[n] = tr;
        tau_d[n] = tr + p.tau_delta_min + (p.tau_delta_max - p.tau_delta_min) * rt.u(2ull * n + 1);
        phi0[n] = p.phi0_lo + (p.phi0_hi - p.phi0_lo) * rp.u(2ull * n);
        phi1[n] = p.phi1_lo + (p.phi1_hi - p.phi1_lo) * rp.u(2ull * n + 1);
        const int rank = (int)conn.inv((uint32_t)n);             // connected = the first n_conn cells of a random order
        double wn = 0.0;
        if (rank < p.n_strong) wn = p.strong_lo + (p.strong_hi - p.strong_lo) * rw.u((unsigned long long)n);
        else if (rank < p.n_conn) wn = -p.weak_mean * log(rw.u((unsigned long long)n)) + p.min_weight;
        w[n] = wn;
        if (weights_out) weights_out[(size_t)b * p.N + n] = wn;
        if (wn != 0.0) { lo = fmin(lo, wn); hi = fmax(hi, wn); }
    }
    __shared__ double slo[256], shi[256];
    slo[threadIdx.x] = lo; shi[threadIdx.x] = hi;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { slo[threadIdx.x] = fmin(slo[threadIdx.x], slo[threadIdx.x + s]); shi[threadIdx.x] = fmax(shi[threadIdx.x], shi[threadIdx.x + s]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { double* wr = (double*)(base + L.wrange); wr[0] = slo[0]; wr[1] = shi[0]; }
}

// Marsaglia-Tsang gamma(alpha >= 1, 1) on a private counter range
__device__ double gamma_mt(const Rng& r, unsigned long long base_idx, double alpha) {
    const double d = alpha - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (unsigned long long it = 0; it < 64; ++it) {
        const double x = r.normal(base_idx * 64 + it);
        const double v1 = 1.0 + c * x;
        if (v1 <= 0.0) continue;
        const double v = v1 * v1 * v1;
        const double uu = r.u((base_idx * 64 + it) * 2 + 0x8000000000000000ull);
        if (log(uu) < 0.5 * x * x + d - d * v + d * log(v)) return d * v;
    }
    return d;
}

// spikes of the connected neurons (simulation.py:92-114): one warp per (map, neuron); entry e = t * H + slot of the
// stimulation (t = un-shuffled trial index) receives the spike latency (uint16, 0 = no spike) and the PSC amplitude
__global__ void __launch_bounds__(128) sim_spikes_kernel(const Params p, const Lay L, char* ws, const unsigned long long* seeds,
                                                         int* status) {
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int n = blockIdx.x * 4 + wl, bb = blockIdx.y;
    if (n >= p.N) return;
    char* base = ws + (size_t)bb * L.stride;
    const unsigned long long seed = seeds[bb];
    const double wn = ((const double*)(base + L.w))[n];
    if (wn == 0.0) return;                                       // unconnected cells leave no trace in the PSCs
    const double ph0 = ((const double*)(base + L.phi0))[n], ph1 = ((const double*)(base + L.phi1))[n];
    unsigned short* ev_time = (unsigned short*)(base + L.ev_time);
    float* ev_amp = (float*)(base + L.ev_amp);
    const Rng rs(seed, S_SPK), rpad(seed, S_PAD), rl(seed, S_LAT), rm(seed, S_MULT);
    __shared__ unsigned int topkey[4][TOPMAX];
    __shared__ int topent[4][TOPMAX];
    int ntop = 0, nspk_top = 0;
    for (int ps0 = 0; ps0 < p.npass; ps0 += 32) {                // lanes take passes; the order of a pass is a keyed permutation
        const int ps = ps0 + lane;
        int ent[PMAX];
        bool spk[PMAX];
        if (ps < p.npass) {
            const Perm order(seed, S_PERMN + (uint32_t)ps, (uint32_t)p.N);
            const int pos = (int)order.inv((uint32_t)n);
            const int h = pos / p.H, slot = pos - h * p.H;
            for (int a = 0; a < p.P; ++a) {
                const long long t = ((long long)ps * p.P + a) * p.nh + h;
                ent[a] = -1; spk[a] = false;
                if (t < p.K) {
                    const int e = (int)(t * p.H + slot);
                    ent[a] = e;
                    const double fr = 1.0 / (1.0 + exp(-(ph0 * p.powers_desc[a] - ph1)));
                    spk[a] = rs.u((unsigned long long)e) <= fr;
                }
            }
        }
        // top power (a = 0): remember the stimulation for the minimum-rate padding
        const bool has_top = ps < p.npass && ent[0] >= 0;
        const unsigned m = __ballot_sync(0xffffffffu, has_top);
        const int off = ntop + __popc(m & ((1u << lane) - 1u));
        if (has_top) {
            if (off < TOPMAX) {
                topent[wl][off] = spk[0] ? -1 - ent[0] : ent[0];           // negative = already spiking
                uint32_t ka, kb;
                rpad.bits((unsigned long long)ent[0], ka, kb);
                topkey[wl][off] = ka;
            }
        }
        ntop += __popc(m);
        nspk_top += __popc(__ballot_sync(0xffffffffu, has_top && spk[0]));
        // the other powers are final now
        if (ps < p.npass)
            for (int a = 1; a < p.P; ++a)
                if (ent[a] >= 0) {
                    unsigned short tm = 0; float amp = 0.f;
                    if (spk[a]) {
                        const double alpha = 1e4 / (p.powers_desc[a] * p.powers_desc[a]);
                        const double lat = p.min_latency + p.gamma_beta * gamma_mt(rl, (unsigned long long)ent[a], alpha);
                        tm = lat < 65535.0 ? (unsigned short)(int)lat : 65535;
                        amp = (float)(wn * exp(p.mult_noise_log_var * rm.normal((unsigned long long)ent[a])));
                    }
                    ev_time[ent[a]] = tm; ev_amp[ent[a]] = amp;
                }
    }
    __syncwarp();
    if (ntop > TOPMAX) { if (lane == 0) atomicExch(&status[bb], CM_EUNSUPPORTED); return; }
    // pad spikes at the top power up to max_power_min_spike_rate (simulation.py:98-108): req random cells of the
    // non-spiking ones = those with the smallest random keys
    int req = 0;
    if (ntop > 0) {
        const double fr = (double)nspk_top / (double)ntop;
        const double diff = p.max_power_min_spike_rate - fr;
        if (diff > 0.0) req = (int)ceil(diff * (double)ntop);
    }
    for (int i = lane; i < ntop; i += 32) {
        int e = topent[wl][i];
        bool sp = e < 0;
        if (sp) e = -1 - e;
        else if (req > 0) {
            const unsigned int ki = topkey[wl][i];
            int rank = 0;
            for (int j = 0; j < ntop; ++j)
                if (topent[wl][j] >= 0) rank += (topkey[wl][j] < ki) || (topkey[wl][j] == ki && j < i);
            sp = rank < req;
        }
        unsigned short tm = 0; float amp = 0.f;
        if (sp) {
            const double alpha = 1e4 / (p.powers_desc[0] * p.powers_desc[0]);
            const double lat = p.min_latency + p.gamma_beta * gamma_mt(rl, (unsigned long long)e, alpha);
            tm = lat < 65535.0 ? (unsigned short)(int)lat : 65535;
            amp = (float)(wn * exp(p.mult_noise_log_var * rm.normal((unsigned long long)e)));
        }
        ev_time[e] = tm; ev_amp[e] = amp;
    }
}

// closed forms for the bi-exponential kernel ke(u) = exp(-u / td) - exp(-u / tr), u = 0, 1, ...
__device__ __forceinline__ double geo_sum(double tau, int n) {           // sum_{u=0}^{n-1} exp(-u / tau)
    return n <= 0 ? 0.0 : (1.0 - exp(-(double)n / tau)) / (1.0 - exp(-1.0 / tau));
}

// one CTA per (shuffled trial j, map): design column + trace
template <typename TP, typename TS>
__global__ void __launch_bounds__(256) sim_traces_kernel(const Params p, const Lay L, char* ws, const unsigned long long* seeds,
                                                         TS* __restrict__ stim, unsigned char* __restrict__ codes,
                                                         TP* __restrict__ psc) {
    const int j = blockIdx.x, b = blockIdx.y;
    char* base = ws + (size_t)b * L.stride;
    const unsigned long long seed = seeds[b];
    const int T = p.T;
    __shared__ float white[TMAX + 2 * FR];
    __shared__ float filt[2 * FR + 1];
    __shared__ float acc[TMAX];
    __shared__ double evd[64][4];           // per slot: start sample, scale, tau_d, tau_r
    __shared__ double spont[4];             // start, scale, tau_d, tau_r
    const Perm shuffle(seed, S_PERMK, (uint32_t)p.K);
    const long long t = (long long)shuffle.fwd((uint32_t)j);              // un-shuffled trial index
    const int per_pass = p.P * p.nh;
    const int ps = (int)(t / per_pass), rem = (int)(t - (long long)ps * per_pass);
    const int a = rem / p.nh, h = rem - a * p.nh;
    if (threadIdx.x < p.H) {
        const int q = threadIdx.x, pos = h * p.H + q;
        int n = -1;
        double sc = 0.0, st = 0.0, td = 1.0, tr = 1.0;
        if (pos < p.N) {
            const Perm order(seed, S_PERMN + (uint32_t)ps, (uint32_t)p.N);
            n = (int)order.fwd((uint32_t)pos);
            const size_t o = ((size_t)b * p.N + n) * (size_t)p.K + j;
            if (stim) stim[o] = (TS)p.powers_desc[a];
            if (codes) codes[o] = p.code_desc[a];
            const double wn = ((const double*)(base + L.w))[n];
            if (wn != 0.0) {
                const int e = (int)(t * p.H + q);
                const int s = ((const unsigned short*)(base + L.ev_time))[e];
                if (s > 0 && s < T) {
                    td = ((const double*)(base + L.tau_d))[n]; tr = ((const double*)(base + L.tau_r))[n];
                    // unit-area kernel (trapz over T samples + 1e-5), shifted to s, renormalised by its sum + 1e-5 (:17-21,285-289)
                    const double full = geo_sum(td, T) - geo_sum(tr, T);
                    const double last = exp(-(double)(T - 1) / td) - exp(-(double)(T - 1) / tr);
                    const double Z = full - 0.5 * last + 1e-5;                       // ke(0) = 0
                    const double part = (geo_sum(td, T - s) - geo_sum(tr, T - s)) / Z;
                    sc = (double)((const float*)(base + L.ev_amp))[e] / Z / (part + 1e-5);
                    st = (double)s;
                }
            }
        }
        evd[q][0] = st; evd[q][1] = sc; evd[q][2] = td; evd[q][3] = tr;
    }
    if (threadIdx.x == 64) {                 // spontaneous PSC (simulation.py:157-170)
        const Rng rsp(seed, S_SPONT);
        double sc = 0.0, st = 0.0, td = 1.0, tr = 1.0;
        if (rsp.u(8ull * j) <= p.spont_prob) {
            tr = p.tau_r_min + (p.tau_r_max - p.tau_r_min) * rsp.u(8ull * j + 1);
            td = tr + p.tau_delta_min + (p.tau_delta_max - p.tau_delta_min) * rsp.u(8ull * j + 2);
            int s = 1 + (int)(rsp.u(8ull * j + 3) * (double)(T - 1));
            if (s > T - 1) s = T - 1;
            const double* wr = (const double*)(base + L.wrange);
            const double wv = wr[0] + (wr[1] - wr[0]) * rsp.u(8ull * j + 4);
            // kern(t) = ke(t - s) for t > s; trapz over T samples = sum - kern(T-1) / 2
            const int len = T - s;
            const double sum = geo_sum(td, len) - geo_sum(tr, len);
            const double last = exp(-(double)(len - 1) / td) - exp(-(double)(len - 1) / tr);
            sc = (wr[1] >= wr[0]) ? wv / (sum - 0.5 * last + 1e-5) : 0.0;
            st = (double)s;
        }
        spont[0] = st; spont[1] = sc; spont[2] = td; spont[3] = tr;
    }
    // correlated noise: white noise filtered with g(d) ~ exp(-d^2 / l^2), sum g^2 = 1  ->  cov(d) = exp(-d^2 / (2 l^2))
    const Rng rg(seed, S_GP), ri(seed, S_IID);
    const float l2 = (float)(p.gp_lengthscale * p.gp_lengthscale);
    for (int i = threadIdx.x; i <= 2 * FR; i += blockDim.x) {
        const float d = (float)(i - FR);
        filt[i] = __expf(-d * d / l2);
    }
    const int nw = T + 2 * FR;
    for (int i = threadIdx.x; 2 * i < nw; i += blockDim.x) {
        const float2 z = rg.normal2f((unsigned long long)j * 2048ull + (unsigned long long)i);
        white[2 * i] = z.x;
        if (2 * i + 1 < nw) white[2 * i + 1] = z.y;
    }
    __syncthreads();
    float g2 = 0.f;
    for (int i = 0; i <= 2 * FR; ++i) g2 += filt[i] * filt[i];
    const float gnorm = (float)p.gp_scale * rsqrtf(g2);
    for (int tt = threadIdx.x; tt < T; tt += blockDim.x) {
        float s = 0.f;
#pragma unroll 8
        for (int i = 0; i <= 2 * FR; ++i) s = fmaf(filt[i], white[tt + i], s);
        acc[tt] = s * gnorm;
    }
    __syncthreads();
    for (int i = threadIdx.x; 2 * i < T; i += blockDim.x) {
        const float2 z = ri.normal2f((unsigned long long)j * 1024ull + (unsigned long long)i);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int tt = 2 * i + half;
            if (tt >= T) break;
            double v = (double)acc[tt] + p.sigma * (double)(half ? z.y : z.x);
            for (int q = 0; q < p.H; ++q) {
                const double sc = evd[q][1];
                if (sc != 0.0) {
                    const double u = (double)tt - evd[q][0];
                    if (u >= 0.0) v += sc * (exp(-u / evd[q][2]) - exp(-u / evd[q][3]));
                }
            }
            if (spont[1] != 0.0) {
                const double u = (double)tt - spont[0];
                if (u > 0.0) v += spont[1] * (exp(-u / spont[2]) - exp(-u / spont[3]));
            }
            psc[((size_t)b * p.K + j) * (size_t)T + tt] = (TP)v;
        }
    }
}

}  // namespace sim
}  // namespace cm

using namespace cm;

extern "C" size_t cm_simulate_workspace_bytes(int B, int N, int K, int H) {
    if (B <= 0 || N <= 0 || K <= 0 || H <= 0) return 0;
    return sim::make_lay(N, K, H).stride * (size_t)B + (size_t)B * 8 + 256;
}

extern "C" int cm_simulate(const cm_sim_options* o, int B, const uint64_t* seeds, void* stim_dev, int stim_dtype,
                           unsigned char* codes_dev, void* psc_dev, int psc_dtype, double* weights_dev, int* status_dev,
                           void* workspace_dev, size_t workspace_bytes, void* stream) {
    reset_launch_count();
    if (!o || !seeds || B <= 0 || !psc_dev || !status_dev || !workspace_dev) { set_error("cm_simulate: null argument"); return CM_EINVAL; }
    if (o->N <= 0 || o->K <= 0 || o->H <= 0 || o->H > 64 || o->T < 2 || o->T > sim::TMAX || o->n_powers < 1 || o->n_powers > sim::PMAX) {
        set_error("cm_simulate: unsupported shape N=%d K=%d T=%d H=%d P=%d (T <= %d, H <= 64, P <= %d)", o->N, o->K, o->T, o->H,
                  o->n_powers, sim::TMAX, sim::PMAX);
        return CM_ESHAPE;
    }
    if ((long long)o->K * o->H > 0x7fffffffll) { set_error("cm_simulate: K * H too large"); return CM_ESHAPE; }
    if (!stim_dev && !codes_dev) { set_error("cm_simulate: need stim_dev or codes_dev"); return CM_EINVAL; }
    const size_t need = cm_simulate_workspace_bytes(B, o->N, o->K, o->H);
    if (workspace_bytes < need) { set_error("cm_simulate: workspace %zu < required %zu bytes", workspace_bytes, need); return CM_EWORKSPACE; }
    sim::Params p{};
    p.N = o->N; p.K = o->K; p.T = o->T; p.H = o->H; p.P = o->n_powers;
    p.nh = (o->N + o->H - 1) / o->H;
    p.npass = (int)(((long long)o->K + (long long)p.P * p.nh - 1) / ((long long)p.P * p.nh));
    if (p.npass > 60000) { set_error("cm_simulate: too many design passes"); return CM_ESHAPE; }
    p.n_conn = (int)(o->connection_prob * o->N);
    p.n_strong = (int)std::ceil(o->frac_strongly_connected * p.n_conn);
    for (int a = 0; a < p.P; ++a) {
        if (a > 0 && !(o->powers[a] > o->powers[a - 1])) { set_error("cm_simulate: powers must be ascending and distinct"); return CM_EINVAL; }
        p.powers_desc[a] = o->powers[p.P - 1 - a];
        p.code_desc[a] = (unsigned char)(p.P - a);
    }
    if (!(o->powers[0] > 0.0)) { set_error("cm_simulate: powers must be positive"); return CM_EINVAL; }
    p.min_latency = o->min_latency; p.gamma_beta = o->gamma_beta; p.sigma = o->sigma;
    p.strong_lo = o->strong_weight_lower; p.strong_hi = o->strong_weight_upper; p.weak_mean = o->weak_exp_mean; p.min_weight = o->min_weight;
    p.phi0_lo = o->phi_0_lower; p.phi0_hi = o->phi_0_upper; p.phi1_lo = o->phi_1_lower; p.phi1_hi = o->phi_1_upper;
    p.mult_noise_log_var = o->mult_noise_log_var; p.tau_r_min = o->tau_r_min; p.tau_r_max = o->tau_r_max;
    p.tau_delta_min = o->tau_delta_min; p.tau_delta_max = o->tau_delta_max; p.gp_scale = o->gp_scale; p.gp_lengthscale = o->gp_lengthscale;
    p.spont_prob = o->spont_prob; p.max_power_min_spike_rate = o->max_power_min_spike_rate;
    for (int a = 0; a < p.P; ++a)
        if (1e4 / (p.powers_desc[a] * p.powers_desc[a]) < 1.0) { set_error("cm_simulate: gamma shape 1e4 / power^2 < 1 unsupported (power > 100)"); return CM_EUNSUPPORTED; }
    const sim::Lay L = sim::make_lay(o->N, o->K, o->H);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace_dev;
    unsigned long long* seeds_dev = (unsigned long long*)(ws + L.stride * (size_t)B);
    CM_CUDA_CHECK(cudaMemcpyAsync(seeds_dev, seeds, (size_t)B * 8, cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(cudaMemsetAsync(status_dev, 0, (size_t)B * sizeof(int), st));
    const size_t nk = (size_t)B * o->N * o->K;
    if (stim_dev) {
        if (stim_dtype != CM_F32 && stim_dtype != CM_F64) { set_error("cm_simulate: stim dtype must be CM_F32 / CM_F64 (codes go to codes_dev)"); return CM_EINVAL; }
        CM_CUDA_CHECK(cudaMemsetAsync(stim_dev, 0, nk * (stim_dtype == CM_F32 ? 4 : 8), st));
    }
    if (codes_dev) CM_CUDA_CHECK(cudaMemsetAsync(codes_dev, 0, nk, st));
    sim::sim_neurons_kernel<<<B, 256, 0, st>>>(p, L, ws, seeds_dev, weights_dev);
    sim::sim_spikes_kernel<<<dim3((o->N + 3) / 4, B), 128, 0, st>>>(p, L, ws, seeds_dev, status_dev);
    const dim3 grid(o->K, B);
#define CM_SIM_LAUNCH(TP, TS) sim::sim_traces_kernel<TP, TS><<<grid, 256, 0, st>>>(p, L, ws, seeds_dev, (TS*)stim_dev, codes_dev, (TP*)psc_dev)
    if (psc_dtype == CM_F32) { if (stim_dtype == CM_F64 && stim_dev) CM_SIM_LAUNCH(float, double); else CM_SIM_LAUNCH(float, float); }
    else if (psc_dtype == CM_F64) { if (stim_dtype == CM_F64 && stim_dev) CM_SIM_LAUNCH(double, double); else CM_SIM_LAUNCH(double, float); }
    else { set_error("cm_simulate: bad psc dtype"); return CM_EINVAL; }
#undef CM_SIM_LAUNCH
    count_launch(3);
    CM_CUDA_CHECK(cudaGetLastError());
    return CM_OK;
}
