/* circuitmap_b200 -- C ABI of the B200-native hot paths of marcustriplett/circuitmap.
 *
 * The reference has no FFI of its own: its boundary is the Python call surface
 *   circuitmap.NeuralDemixer(path)(traces)                   circuitmap/neural_waveform_demixing.py:17-54
 *   circuitmap.Model(N).fit(psc, stim, 'caviar', fit_options) circuitmap/model.py:36-44,104-162
 *                                      -> optimise.caviar(...) circuitmap/optimise/caviar.py:20-100
 * circuitmap_b200/{neural_waveform_demixing,model}.py keep that surface and marshal to the entry
 * points below (ctypes; see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - plain C types only; every *_dev pointer is DEVICE memory owned by the caller;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are
 *     stream-ordered and return without synchronising unless stated otherwise;
 *   - return 0 on success, a CM_E* code otherwise; cm_last_error() gives the thread-local text;
 *   - re-entrant: no global state besides the per-handle weights.
 */
#ifndef CIRCUITMAP_B200_H
#define CIRCUITMAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CM_VERSION 100
#if defined(__GNUC__)
#define CM_API __attribute__((visibility("default")))
#else
#define CM_API
#endif

enum { CM_OK = 0, CM_EINVAL = 1, CM_ESHAPE = 2, CM_EUNSUPPORTED = 3, CM_ECUDA = 4, CM_EWORKSPACE = 5 };
enum { CM_F32 = 0, CM_F64 = 1, CM_U8 = 2 };   /* CM_U8: stimulus designs only, see cm_caviar_args.stim_dev */

CM_API int cm_version(void);
CM_API const char* cm_last_error(void);
/* number of SMs / device ordinal of the current CUDA device (diagnostics for the host layer) */
CM_API int cm_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * NWD demixer forward  (replaces NeuralDemixer.__call__ / NWDUNet.forward,
 *                       neural_waveform_demixing.py:36-54,204-287,337-348)
 * ------------------------------------------------------------------------------------------ */
typedef struct cm_nwd cm_nwd_t;

/* CM_NWD_NUM_TENSORS host float32 tensors in state_dict order (neural_waveform_demixing.py:259-269):
 * for each of dblock1..4, ublock1..4, conv: {conv|deconv}.weight, .bias, bn.weight, bn.bias,
 * bn.running_mean, bn.running_var.  BatchNorm (eval) is folded into the convolutions inside. */
#define CM_NWD_NUM_TENSORS 54
#define CM_NWD_T 900
CM_API int  cm_nwd_create(const float* const* tensors, int n_tensors, cm_nwd_t** out);
CM_API void cm_nwd_destroy(cm_nwd_t* h);
/* arithmetic of the convolution layers: 0 = fp32 on CUDA cores (default; same arithmetic as the reference's fp32
 * network); 1 = TF32 operands / fp32 accumulation on the 5th-gen tensor cores (tcgen05, TMEM accumulators), one trace
 * per CTA (csrc/nwd_tc.cu); 2 = fp16 operands (11-bit significand, as TF32) / fp32 accumulation, all nine
 * convolutions on tcgen05 with several traces per M tile (csrc/nwd_mt.cu) -- the fast path.  Error bound of modes 1
 * and 2 on unit-normalised traces: max-abs <= 2e-2, relative L2 <= 3e-3 (asserted in tests/test_nwd_gpu.py).
 * Mode 2 keeps weights and activations in fp16 and never saturates silently: a trace whose normalised samples
 * |x / max(x)| exceed 6e4, whose max is 0, that holds a non-finite sample or that drives any activation of the network
 * beyond 6e4 yields an all-NaN row; cm_nwd_set_precision(h, 2) fails if a BatchNorm-folded weight exceeds the fp16 range. */
CM_API int  cm_nwd_set_precision(cm_nwd_t* h, int precision);

/* traces_dev: K x T row-major (in_dtype), out_dev: K x T row-major (out_dtype).
 * Per trace: x = trace / max_t(trace); net(x) in fp32; out = net * max; running minimum from
 * column monotone_start (nwd.py:337-343; monotone_start >= T or < 1 disables it).
 * y_dev / ss_dev (optional, may be NULL): per-trace trapz(out) and sum(out^2) in fp64 -- the
 * CAVIaR prologue statistics (caviar.py:28,30), so a following cm_caviar_fit need not re-read
 * the K x T array.  T must be CM_NWD_T. */
CM_API int  cm_nwd_forward(cm_nwd_t* h, const void* traces_dev, int in_dtype, void* out_dev, int out_dtype,
                    int K, int T, int monotone_start, double* y_dev, double* ss_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * CAVIaR fit  (replaces optimise.caviar, caviar.py:20-316 + pava.py:9-88)
 * B independent fits of identical (N, K) per call; B = 1 for Model.fit.
 * ------------------------------------------------------------------------------------------ */
#define CM_CAVIAR_MAX_POWERS 16

typedef struct cm_caviar_options {      /* keyword arguments of caviar(), caviar.py:21-23 */
    int    iters;                /* 50   */
    int    num_mc_samples;       /* 100  */
    double y_xcorr_thresh;       /* 1e-2 */
    double minimum_spike_count;  /* 3    */
    int    delay_spont_est;      /* 1    */
    double msrmp;                /* 0.3  */
    double scale_factor;         /* 0.75 */
    double penalty;              /* 5.0  */
    int    max_backtrack_iters;  /* 20 (soft-threshold passes, caviar.py:86) */
    double tol;                  /* 0.05 */
    double spont_orthogonality;  /* 0.1  */
    int    fn_scan;              /* 1    */
    int    save_histories;       /* 0    */
} cm_caviar_options;

typedef struct cm_caviar_args {
    int B, N, K, T;              /* fits, neurons, trials, samples per trace (T ignored if psc_dev==NULL) */
    /* inputs ------------------------------------------------------------------------------ */
    const void*   psc_dev;       /* B x K x T (psc_dtype) or NULL when y_dev/ss_dev are given      */
    int           psc_dtype;
    const double* y_dev;         /* B x K  trapz(psc)   (used when psc_dev == NULL)                */
    const double* ss_dev;        /* B x K  sum_t psc^2  (used when psc_dev == NULL)                */
    const void*   stim_dev;      /* B x N x K, neuron-major rows, K contiguous (stim_dtype): laser powers (CM_F32 /
                                    CM_F64, 0 = not targeted, README.md:26) or power codes (CM_U8: c in 1..P means
                                    powers[c-1], 0 = not targeted; see cm_pack_stim_u8)                            */
    int           stim_dtype;
    int           n_powers;      /* P = number of distinct non-zero powers (<= CM_CAVIAR_MAX_POWERS) */
    const double* powers;        /* HOST array, ascending: np.unique(stim)[1:] (caviar.py:42)      */
    const uint64_t* seeds;       /* HOST array, B seeds (caviar.py:76)                             */
    /* priors, B x ... fp64 (model.py:24-31) --------------------------------------------------- */
    const double* mu0_dev;       /* B x N     */
    const double* beta0_dev;     /* B x N     */
    const double* phi0_dev;      /* B x N x 2 */
    const double* phi_cov0_dev;  /* B x N x 2 x 2 */
    const double* shape0;        /* HOST, B   */
    const double* rate0;         /* HOST, B   */
    cm_caviar_options opt;
    /* outputs, fp64 --------------------------------------------------------------------------- */
    double* mu_dev;              /* B x N */
    double* beta_dev;            /* B x N */
    double* lam_dev;             /* B x N x K dense, or NULL to skip densification */
    double* shape_dev;           /* B */
    double* rate_dev;            /* B */
    double* phi_dev;             /* B x N x 2 */
    double* phi_cov_dev;         /* B x N x 2 x 2 */
    double* z_dev;               /* B x K */
    /* optional per-iteration histories (opt.save_histories); NULL to skip individual ones ------ */
    double* mu_hist_dev;         /* B x iters x N */
    double* beta_hist_dev;       /* B x iters x N */
    double* lam_hist_dev;        /* B x iters x N x K */
    double* shape_hist_dev;      /* B x iters */
    double* rate_hist_dev;       /* B x iters */
    double* phi_hist_dev;        /* B x iters x N x 2 */
    double* phi_cov_hist_dev;    /* B x iters x N x 2 x 2 */
    double* z_hist_dev;          /* B x iters x K */
    /* workspace -------------------------------------------------------------------------------- */
    int64_t nnz_cap;             /* upper bound on non-zeros of stim PER FIT */
    void*   workspace_dev;
    size_t  workspace_bytes;     /* >= cm_caviar_workspace_bytes(B, N, K, nnz_cap, flags) */
    int*    status_dev;          /* B ints: 0 ok, else CM_E* detected on device (1 invalid stimulus entry, 5 nnz overflow,
                                    9 a helper CTA of a large single fit never answered, 10 the two chain teams of a
                                    sweep lost each other -- both are bounded waits instead of hangs, never seen in practice) */
    /* optional sparse posterior: CSR over (neuron, trial) of the entries lam can be non-zero on, i.e. targeted trials
     * that pass the lam_mask (caviar.py:30-34,216) -- 8 nnz bytes instead of the 8 N K of lam_dev.  All three or none. */
    double*  lam_csr_val_dev;    /* B x nnz_cap */
    int32_t* lam_csr_col_dev;    /* B x nnz_cap  trial index of every entry */
    int32_t* lam_csr_ptr_dev;    /* B x (N + 1)  row pointers; entries past ptr[N] are unspecified */
    /* CTA variant of the persistent kernel: 0 = automatic (512-thread CTAs, one per SM, while the batch fits one wave of them;
     * 256-thread CTAs, two per SM, beyond), 256 / 512 = forced.  Launches that are meant to overlap on the GPU (chunks of a
     * stream, streaming.FitPipeline) force 256 so that CTAs of two launches share an SM. */
    int cta_variant;
} cm_caviar_args;

CM_API size_t cm_caviar_workspace_bytes(int B, int N, int K, int64_t nnz_cap, int save_histories);
CM_API int    cm_caviar_fit(const cm_caviar_args* a, void* stream);

/* One streaming pass over a dense device-resident design (count = B*N*K entries of `dtype`): number of non-zero entries and
 * the sorted distinct non-zero values (what the reference derives on the host with np.unique, caviar.py:42).
 * values_out: HOST array of CM_CAVIAR_MAX_POWERS + 2 doubles; *n_values_out > CM_CAVIAR_MAX_POWERS + 1 means "too many".
 * scratch_dev: cm_caviar_scan_scratch_bytes() bytes of device memory.  Synchronises the stream. */
CM_API size_t cm_caviar_scan_scratch_bytes(void);
CM_API int    cm_caviar_scan_stim(const void* stim_dev, int dtype, int64_t count, void* scratch_dev, int64_t* nnz_out,
                                  double* values_out, int* n_values_out, void* stream);
/* Sparse design -> dense uint8 codes ON THE DEVICE: nnz (neuron, trial, code) triples (device arrays) are scattered into the zeroed
 * N x K code matrix codes_out_dev that cm_caviar_fit takes as a CM_U8 stimulus.  A compressive design has nnz <= K H entries, so
 * 9 nnz bytes cross the bus instead of N K.  status_dev (one int, may be NULL) receives CM_EINVAL for an out-of-range index. */
CM_API int    cm_expand_stim_coo(const int* neuron_dev, const int* trial_dev, const unsigned char* code_dev, int64_t nnz, int N, int K,
                                 unsigned char* codes_out_dev, int* status_dev, void* stream);
/* HOST helper (no device work, `threads` worker threads): powers = np.unique(stim)[1:] (caviar.py:42) into powers_out
 * (CM_CAVIAR_MAX_POWERS doubles), the number of non-zero entries, and -- if codes_out != NULL -- the design as uint8
 * power codes for the CM_U8 stimulus dtype (N*K bytes to upload instead of 8*N*K).  stim_host: CM_F32 or CM_F64. */
CM_API int    cm_pack_stim_u8(const void* stim_host, int dtype, int64_t count, double* powers_out, int* n_powers_out,
                              int64_t* nnz_out, unsigned char* codes_out, int threads);

/* ------------------------------------------------------------------------------------------
 * Synthetic mapping experiments on the device  (circuitmap/simulation.py:25-215, blockwise design, nreps = 1).
 * The reference draws from NumPy's unseeded global stream: parity is at the level of the distributions
 * (tests/test_simulate_gpu.py), not of individual draws.  B maps per call, one seed each.
 * ------------------------------------------------------------------------------------------ */
typedef struct cm_sim_options {          /* keyword arguments of simulate(), simulation.py:25-29 (defaults in brackets) */
    int    N, K, T, H;                   /* neurons [300], trials [1000], samples per trace [900], targets per hologram [10] */
    int    n_powers;                     /* [3] */
    double powers[CM_CAVIAR_MAX_POWERS]; /* ascending [45, 55, 65]; each <= 100 (gamma shape 1e4 / power^2 >= 1) */
    double connection_prob;              /* [0.05] */
    double frac_strongly_connected;      /* [0.2] */
    double min_latency, gamma_beta;      /* [160, 15] spike latency = min_latency + Gamma(1e4 / power^2, gamma_beta) */
    double sigma;                        /* [6e-4] iid noise */
    double strong_weight_lower, strong_weight_upper, weak_exp_mean, min_weight;          /* [20, 40, 4, 9] */
    double phi_0_lower, phi_0_upper, phi_1_lower, phi_1_upper;                           /* [0.2, 0.25, 10, 15] */
    double mult_noise_log_var;           /* [0.01] */
    double tau_r_min, tau_r_max, tau_delta_min, tau_delta_max;                           /* [25, 60, 75, 250] */
    double gp_scale, gp_lengthscale;     /* [4e-3, 50] */
    double spont_prob;                   /* [0.05] */
    double max_power_min_spike_rate;     /* [0.4] */
} cm_sim_options;

CM_API size_t cm_simulate_workspace_bytes(int B, int N, int K, int H);
/* stim_dev: B x N x K laser powers (CM_F32 / CM_F64) or NULL; codes_dev: B x N x K uint8 power codes or NULL (at least one);
 * psc_dev: B x K x T traces (CM_F32 / CM_F64); weights_dev: B x N ground-truth synaptic weights (may be NULL);
 * status_dev: B ints (0 ok, CM_EUNSUPPORTED if a neuron has more than 1024 top-power trials); seeds: HOST array of B. */
CM_API int cm_simulate(const cm_sim_options* o, int B, const uint64_t* seeds, void* stim_dev, int stim_dtype,
                       unsigned char* codes_dev, void* psc_dev, int psc_dtype, double* weights_dev, int* status_dev,
                       void* workspace_dev, size_t workspace_bytes, void* stream);

/* diagnostics: per-phase SM cycle counters of fit 0 of the persistent kernel (see csrc/caviar.cu phase_mark ids);
 * copies up to n counters to `out` (may be NULL), then clears them and sets the enable flag (synchronises). */
CM_API int cm_caviar_debug_phase_cycles(long long* out, int n, int enable);

/* diagnostics: per-stage SM cycle counters of CTA 0 of the tensor-core demixer kernel (csrc/nwd_tc.cu NWD_MARK ids) */
CM_API int cm_nwd_debug_cycles(long long* out, int n, int enable);
/* same for the multi-trace kernel (csrc/nwd_mt.cu MT_MARK ids) */
CM_API int cm_nwd_mt_debug_cycles(long long* out, int n, int enable);
/* diagnostics: CTA 0 of the next multi-trace forward copies its shared memory (>= 216 KB) to dev_buf when it reaches
 * MT_MARK `stage` of its first pass (stage < 0 disables) */
CM_API int cm_nwd_mt_debug_dump(void* dev_buf, int stage);
/* test hook (host only, no device work): the packed fp16 tap tables + fp32 biases the multi-trace kernel streams
 * (layout: csrc/nwd_mt.cuh).  out may be NULL to query the size. */
CM_API int cm_nwd_mt_pack(const float* const* tensors, int n_tensors, unsigned char* out, size_t cap, size_t* need);

/* number of kernel launches issued by the last cm_* call on this thread (for bench accounting) */
CM_API int cm_last_launch_count(void);
/* device time (ms, CUDA events on the caller's stream) of the dominant kernel of the last cm_nwd_forward /
 * cm_caviar_fit call on this thread: the U-Net kernel, resp. the persistent fit kernel.  Synchronises on the
 * kernel's end event.  Returns a negative value if no timed kernel was launched. */
CM_API float cm_last_main_kernel_ms(void);

#ifdef __cplusplus
}
#endif
#endif
