"""ctypes binding of libcircuitmap_b200.so (the C ABI declared in include/circuitmap_b200.h).

There is NO CPU fallback: if the library is missing or CUDA is unavailable the import of the hot
paths fails loudly.  Build with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C circuitmap_b200/csrc`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcircuitmap_b200.so")

CM_F32, CM_F64, CM_U8 = 0, 1, 2
CM_OK, CM_EINVAL, CM_ESHAPE, CM_EUNSUPPORTED, CM_ECUDA, CM_EWORKSPACE = 0, 1, 2, 3, 4, 5
CM_NWD_NUM_TENSORS = 54
CM_NWD_T = 900
CM_CAVIAR_MAX_POWERS = 16


class CaviarOptions(C.Structure):
    _fields_ = [("iters", C.c_int), ("num_mc_samples", C.c_int), ("y_xcorr_thresh", C.c_double),
                ("minimum_spike_count", C.c_double), ("delay_spont_est", C.c_int), ("msrmp", C.c_double),
                ("scale_factor", C.c_double), ("penalty", C.c_double), ("max_backtrack_iters", C.c_int),
                ("tol", C.c_double), ("spont_orthogonality", C.c_double), ("fn_scan", C.c_int),
                ("save_histories", C.c_int)]


class CaviarArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("N", C.c_int), ("K", C.c_int), ("T", C.c_int),
                ("psc_dev", C.c_void_p), ("psc_dtype", C.c_int),
                ("y_dev", C.c_void_p), ("ss_dev", C.c_void_p),
                ("stim_dev", C.c_void_p), ("stim_dtype", C.c_int),
                ("n_powers", C.c_int), ("powers", C.POINTER(C.c_double)), ("seeds", C.POINTER(C.c_uint64)),
                ("mu0_dev", C.c_void_p), ("beta0_dev", C.c_void_p), ("phi0_dev", C.c_void_p),
                ("phi_cov0_dev", C.c_void_p), ("shape0", C.POINTER(C.c_double)), ("rate0", C.POINTER(C.c_double)),
                ("opt", CaviarOptions),
                ("mu_dev", C.c_void_p), ("beta_dev", C.c_void_p), ("lam_dev", C.c_void_p),
                ("shape_dev", C.c_void_p), ("rate_dev", C.c_void_p), ("phi_dev", C.c_void_p),
                ("phi_cov_dev", C.c_void_p), ("z_dev", C.c_void_p),
                ("mu_hist_dev", C.c_void_p), ("beta_hist_dev", C.c_void_p), ("lam_hist_dev", C.c_void_p),
                ("shape_hist_dev", C.c_void_p), ("rate_hist_dev", C.c_void_p), ("phi_hist_dev", C.c_void_p),
                ("phi_cov_hist_dev", C.c_void_p), ("z_hist_dev", C.c_void_p),
                ("nnz_cap", C.c_int64), ("workspace_dev", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("status_dev", C.c_void_p),
                ("lam_csr_val_dev", C.c_void_p), ("lam_csr_col_dev", C.c_void_p), ("lam_csr_ptr_dev", C.c_void_p),
                ("cta_variant", C.c_int)]


class SimOptions(C.Structure):
    _fields_ = [("N", C.c_int), ("K", C.c_int), ("T", C.c_int), ("H", C.c_int), ("n_powers", C.c_int),
                ("powers", C.c_double * CM_CAVIAR_MAX_POWERS)] + \
               [(k, C.c_double) for k in ("connection_prob", "frac_strongly_connected", "min_latency", "gamma_beta", "sigma",
                                          "strong_weight_lower", "strong_weight_upper", "weak_exp_mean", "min_weight",
                                          "phi_0_lower", "phi_0_upper", "phi_1_lower", "phi_1_upper", "mult_noise_log_var",
                                          "tau_r_min", "tau_r_max", "tau_delta_min", "tau_delta_max", "gp_scale",
                                          "gp_lengthscale", "spont_prob", "max_power_min_spike_rate")]


# every symbol include/circuitmap_b200.h declares
EXPORTS = ["cm_version", "cm_last_error", "cm_device_info", "cm_nwd_create", "cm_nwd_destroy", "cm_nwd_forward",
           "cm_nwd_set_precision",
           "cm_caviar_workspace_bytes", "cm_caviar_fit", "cm_caviar_scan_stim", "cm_caviar_scan_scratch_bytes",
           "cm_pack_stim_u8", "cm_expand_stim_coo", "cm_simulate", "cm_simulate_workspace_bytes", "cm_last_launch_count", "cm_last_main_kernel_ms",
           "cm_caviar_debug_phase_cycles", "cm_nwd_debug_cycles", "cm_nwd_mt_debug_cycles", "cm_nwd_mt_debug_dump", "cm_nwd_mt_pack"]

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"circuitmap_b200: native library {LIB_PATH} is missing. There is no CPU fallback; build it with "
            "`make -C circuitmap_b200/csrc` (needs nvcc with sm_100a support).")
    lib = C.CDLL(LIB_PATH)
    lib.cm_version.restype = C.c_int
    lib.cm_last_error.restype = C.c_char_p
    lib.cm_last_launch_count.restype = C.c_int
    lib.cm_last_main_kernel_ms.restype = C.c_float
    lib.cm_device_info.argtypes = [C.POINTER(C.c_int)] * 3
    lib.cm_nwd_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
    lib.cm_nwd_destroy.argtypes = [C.c_void_p]
    lib.cm_nwd_destroy.restype = None
    lib.cm_nwd_set_precision.argtypes = [C.c_void_p, C.c_int]
    lib.cm_nwd_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cm_caviar_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int]
    lib.cm_caviar_workspace_bytes.restype = C.c_size_t
    lib.cm_caviar_fit.argtypes = [C.POINTER(CaviarArgs), C.c_void_p]
    lib.cm_caviar_scan_scratch_bytes.restype = C.c_size_t
    lib.cm_caviar_scan_stim.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.POINTER(C.c_int64),
                                        C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p]
    lib.cm_expand_stim_coo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    lib.cm_simulate_workspace_bytes.restype = C.c_size_t
    lib.cm_simulate_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    lib.cm_simulate.argtypes = [C.POINTER(SimOptions), C.c_int, C.POINTER(C.c_uint64), C.c_void_p, C.c_int, C.c_void_p,
                                C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.cm_pack_stim_u8.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int64), C.c_void_p, C.c_int]
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().cm_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("circuitmap_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback.")
    return torch
