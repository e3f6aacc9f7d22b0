"""`simulate(...)` with the reference's signature (circuitmap/simulation.py:25-29) on the device (csrc/simulate.cu).

The reference generator draws from NumPy's global Mersenne-Twister stream without seeding it, so no output exists to be
identical to: this one draws from the same distributions with a counter-based stream keyed by `seed` (an extension of the
signature) and is tested on moments and design invariants.  Supported: design='blockwise', nreps=1, sampled weights /
kernels / sigmoid coefficients (the defaults of every experiment script); anything else raises NotImplementedError.
`simulate_batch` generates B maps in one call and leaves them on the device in the compact formats the fit consumes
(float32 traces, uint8 power codes) -- what the sweeps of scripts/run_simulations.py need.
"""
import ctypes as C

import numpy as np

from . import _lib

_DEFAULTS = dict(N=300, T=900, H=10, trials=1000, nreps=1, connection_prob=0.05, powers=[45, 55, 65], min_latency=160,
                 gamma_beta=1.5e1, sigma=6e-4, frac_strongly_connected=0.2, strong_weight_lower=20, strong_weight_upper=40,
                 weak_exp_mean=4, min_weight=9, phi_0_lower=0.2, phi_0_upper=0.25, phi_1_lower=10, phi_1_upper=15,
                 mult_noise_log_var=0.01, tau_r_min=25, tau_r_max=60, tau_delta_min=75, tau_delta_max=250, weights=None,
                 kernel=None, phi_0=None, phi_1=None, gp_scale=4e-3, gp_lengthscale=50, spont_prob=0.05, design="blockwise",
                 max_power_min_spike_rate=0.4, batch_size=500, neuron_batch_size=500)


def _options(kw):
    if kw["design"] != "blockwise":
        assert kw["design"] in ["random", "blockwise"]                    # simulation.py:31
        raise NotImplementedError("device simulate: only design='blockwise' is implemented")
    if kw["nreps"] != 1:
        raise NotImplementedError("device simulate: only nreps=1 is implemented")
    for k in ("weights", "kernel", "phi_0", "phi_1"):
        if kw[k] is not None:
            raise NotImplementedError("device simulate: %s must be None (sampled on the device)" % k)
    o = _lib.SimOptions()
    o.N, o.K, o.T, o.H = int(kw["N"]), int(kw["trials"]), int(kw["T"]), int(kw["H"])
    powers = np.sort(np.asarray(kw["powers"], dtype=np.float64))
    o.n_powers = powers.size
    if not 1 <= powers.size <= _lib.CM_CAVIAR_MAX_POWERS:
        raise ValueError("device simulate: 1..%d powers" % _lib.CM_CAVIAR_MAX_POWERS)
    for i, p in enumerate(powers):
        o.powers[i] = float(p)
    for k in ("connection_prob", "frac_strongly_connected", "min_latency", "gamma_beta", "sigma", "strong_weight_lower",
              "strong_weight_upper", "weak_exp_mean", "min_weight", "phi_0_lower", "phi_0_upper", "phi_1_lower", "phi_1_upper",
              "mult_noise_log_var", "tau_r_min", "tau_r_max", "tau_delta_min", "tau_delta_max", "gp_scale", "gp_lengthscale",
              "spont_prob", "max_power_min_spike_rate"):
        setattr(o, k, float(kw[k]))
    return o, powers


def simulate_batch(seeds, device=None, psc_dtype=None, want_stim=False, workspace=None, out=None, **kwargs):
    """B synthetic maps on the device: returns dict(psc (B, K, T) float32 [or psc_dtype], codes (B, N, K) uint8 power codes,
    weights (B, N) float64, powers (ascending ndarray)[, stim_matrix (B, N, K) float64 when want_stim]).  `out` may carry the
    tensors of an earlier call with the same shapes (reused)."""
    torch = _lib.require_cuda()
    unknown = set(kwargs) - set(_DEFAULTS)
    if unknown:
        raise TypeError("simulate() got an unexpected keyword argument %r" % sorted(unknown)[0])
    kw = dict(_DEFAULTS)
    kw.update(kwargs)
    o, powers = _options(kw)
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    B, N, K, T, H = len(seeds), o.N, o.K, o.T, o.H
    psc_dtype = psc_dtype or torch.float32
    reuse = out or {}

    def buf(name, shape, dtype):
        t = reuse.get(name)
        if t is not None and tuple(t.shape) == tuple(shape) and t.dtype == dtype and t.device == dev:
            return t
        return torch.empty(shape, dtype=dtype, device=dev)

    res = dict(psc=buf("psc", (B, K, T), psc_dtype), codes=buf("codes", (B, N, K), torch.uint8),
               weights=buf("weights", (B, N), torch.float64), status=buf("status", (B,), torch.int32), powers=powers)
    if want_stim:
        res["stim_matrix"] = buf("stim_matrix", (B, N, K), torch.float64)
    need = int(lib.cm_simulate_workspace_bytes(B, N, K, H))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    seeds_arr = (C.c_uint64 * B)(*[int(s) & 0xFFFFFFFFFFFFFFFF for s in seeds])
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib.cm_simulate(C.byref(o), B, seeds_arr, res["stim_matrix"].data_ptr() if want_stim else None, _lib.CM_F64,
                             res["codes"].data_ptr(), res["psc"].data_ptr(),
                             {torch.float32: _lib.CM_F32, torch.float64: _lib.CM_F64}[psc_dtype],
                             res["weights"].data_ptr(), res["status"].data_ptr(), workspace.data_ptr(), workspace.numel(),
                             C.c_void_p(stream))
    _lib.check(rc, "cm_simulate")
    res["_workspace"] = workspace
    res["launches"] = lib.cm_last_launch_count()
    return res


def simulate(N=300, T=900, H=10, trials=1000, seed=0, device=None, **kwargs):
    """Drop-in for circuitmap.simulate (simulation.py:25-195): returns the `sim` dict with NumPy arrays
    'weights' (N,), 'stim_matrix' (N, K) float64, 'psc' (K, T) float64, 'I' (K,) -- the keys the scripts and the fit consume
    (simulation.py:178-192; 'spks', 'spk_times', 'mult_noise', 'gp_noise', 'spont_pscs' are not materialised)."""
    print("Creating simulation with specifications:")
    for k, v in (("N", N), ("T", T), ("H", H), ("Trials", trials)):
        print(k, v)
    torch = _lib.require_cuda()
    res = simulate_batch([seed], device=device, psc_dtype=torch.float64, want_stim=True, N=N, T=T, H=H, trials=trials, **kwargs)
    if int(res["status"].sum().item()) != 0:
        raise RuntimeError("cm_simulate: a neuron has more top-power trials than the generator supports")
    stim = res["stim_matrix"][0].cpu().numpy()
    print("Complete.\n")
    return {"weights": res["weights"][0].cpu().numpy(), "stim_matrix": stim, "psc": res["psc"][0].cpu().numpy(),
            "I": stim.max(0), "sigma": float(kwargs.get("sigma", _DEFAULTS["sigma"]))}
