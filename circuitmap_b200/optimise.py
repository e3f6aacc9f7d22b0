"""`caviar(...)` with the reference's signature (circuitmap/optimise/caviar.py:20-23,100) on the B200 kernel.

Host side only marshals: the traces go to the device as they are, the stimulus design as uint8 power codes (packed by
the threaded C helper `cm_pack_stim_u8`), `cm_caviar_fit` (csrc/caviar.cu) does every arithmetic step of the fit, and
the 17-tuple the reference returns is rebuilt from its outputs.  CUDA tensors are accepted for both inputs (e.g. the
output of `NeuralDemixer` on a device tensor, which carries the y / sum-of-squares hand-off).
`caviar_batched` is the same call for B independent maps of identical (N, K) -- the unit that is sharded over
GPUs (simulation sweeps, LOHO-CV folds).
"""
import ctypes as C

import numpy as np

from . import _lib, _staging

_DEFAULTS = dict(iters=50, num_mc_samples=100, seed=0, y_xcorr_thresh=1e-2, minimum_spike_count=3, delay_spont_est=1,
                 msrmp=0.3, scale_factor=0.75, penalty=5e0, save_histories=False, max_backtrack_iters=20, tol=0.05,
                 spont_orthogonality=0.1, fn_scan=True)


def _options(kw):
    o = _lib.CaviarOptions()
    o.iters = int(kw["iters"])
    o.num_mc_samples = int(kw["num_mc_samples"])
    o.y_xcorr_thresh = float(kw["y_xcorr_thresh"])
    o.minimum_spike_count = float(kw["minimum_spike_count"])
    o.delay_spont_est = int(kw["delay_spont_est"])
    o.msrmp = float(kw["msrmp"])
    o.scale_factor = float(kw["scale_factor"])
    o.penalty = float(kw["penalty"])
    o.max_backtrack_iters = int(kw["max_backtrack_iters"])
    o.tol = float(kw["tol"])
    o.spont_orthogonality = float(kw["spont_orthogonality"])
    o.fn_scan = 1 if kw["fn_scan"] else 0
    o.save_histories = 1 if kw["save_histories"] else 0
    return o


def _validate_options(fit_options):
    """Unknown keyword -> TypeError, exactly like splatting into caviar(**fit_options) (model.py:107-110)."""
    unknown = set(fit_options) - set(_DEFAULTS)
    if unknown:
        raise TypeError("caviar() got an unexpected keyword argument %r" % sorted(unknown)[0])
    kw = dict(_DEFAULTS)
    kw.update(fit_options)
    return kw


def _dt(t):
    import torch
    return {torch.float32: _lib.CM_F32, torch.float64: _lib.CM_F64, torch.uint8: _lib.CM_U8}[t.dtype]


def scan_stim(stim_dev):
    """(nnz, sorted distinct non-zero values) of a dense device-resident design in one streaming pass of
    `cm_caviar_scan_stim` (csrc/stim.cu) -- what the reference gets from np.unique on the host (caviar.py:42)."""
    import torch
    lib = _lib.load()
    stim_dev = stim_dev.contiguous()
    scratch = torch.empty(int(lib.cm_caviar_scan_scratch_bytes()), dtype=torch.uint8, device=stim_dev.device)
    nnz, nvals = C.c_int64(), C.c_int()
    vals = (C.c_double * (_lib.CM_CAVIAR_MAX_POWERS + 2))()
    with torch.cuda.device(stim_dev.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.cm_caviar_scan_stim(stim_dev.data_ptr(), _dt(stim_dev), stim_dev.numel(), scratch.data_ptr(),
                                           C.byref(nnz), vals, C.byref(nvals), C.c_void_p(stream)), "cm_caviar_scan_stim")
    n = nvals.value
    if n > _lib.CM_CAVIAR_MAX_POWERS + 1:
        raise RuntimeError("cm_caviar_fit supports 1..%d distinct stimulus powers, got more" % _lib.CM_CAVIAR_MAX_POWERS)
    return int(nnz.value), np.array(vals[:n], dtype=np.float64)


def powers_like_reference(nnz, values, total):
    """np.unique(I)[1:] (caviar.py:42) from the scan: the sorted distinct values without the smallest one (0 when the
    design holds a zero, which every real design does; quirk A.3 #8 is kept for one that does not)."""
    vals = np.sort(np.concatenate([values, [0.0]])) if nnz < total else np.sort(values)
    return vals[1:]


_pinned_codes = {}


def pack_stim_host(I, threads=None):
    """Host float design (N, K), C-contiguous -> (pinned uint8 code tensor (N, K), powers, nnz) with the threaded C
    helper `cm_pack_stim_u8`: the reference-facing call then uploads N*K bytes instead of 8*N*K."""
    import os
    import torch
    lib = _lib.load()
    dt = {np.dtype(np.float32): _lib.CM_F32, np.dtype(np.float64): _lib.CM_F64}[I.dtype]
    n = I.size
    buf = _pinned_codes.get("buf")
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 1 << 20), dtype=torch.uint8)
        if torch.cuda.is_available():                     # (the helper itself is host-only and is unit-tested without a GPU)
            buf = buf.pin_memory()
        _pinned_codes["buf"] = buf
    powers = (C.c_double * _lib.CM_CAVIAR_MAX_POWERS)()
    P, nnz = C.c_int(), C.c_int64()
    threads = threads or max(1, min(16, (os.cpu_count() or 2) // 2))
    rc = lib.cm_pack_stim_u8(I.ctypes.data_as(C.c_void_p), dt, n, powers, C.byref(P), C.byref(nnz),
                             C.c_void_p(buf.data_ptr()), threads)
    if rc == _lib.CM_EUNSUPPORTED:
        raise RuntimeError("cm_caviar_fit supports 1..%d distinct stimulus powers, got %d"
                           % (_lib.CM_CAVIAR_MAX_POWERS, P.value))
    _lib.check(rc, "cm_pack_stim_u8")
    return buf[:n].view(I.shape), np.array(powers[:P.value], dtype=np.float64), int(nnz.value)


def codes_to_coo(codes):
    """Dense uint8 design codes (N, K) (host or device tensor) -> pinned host triples (neuron int32, trial int32, code uint8),
    ordered by neuron then trial.  Set-up helper for the sparse upload of `streaming.FitPipeline(design='coo')`."""
    import torch
    idx = torch.nonzero(codes)                       # (nnz, 2), row-major order
    vals = codes[idx[:, 0], idx[:, 1]]
    pin = (lambda t: t.cpu().pin_memory()) if torch.cuda.is_available() else (lambda t: t.cpu())
    return pin(idx[:, 0].to(torch.int32).contiguous()), pin(idx[:, 1].to(torch.int32).contiguous()), pin(vals.contiguous())


def expand_coo(neuron_dev, trial_dev, code_dev, N, K, out=None):
    """Device triples -> dense uint8 code matrix (N, K) on the device (`cm_expand_stim_coo`, csrc/stim.cu)."""
    import torch
    lib = _lib.load()
    dev = code_dev.device
    if out is None:
        out = torch.empty((N, K), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.cm_expand_stim_coo(neuron_dev.data_ptr(), trial_dev.data_ptr(), code_dev.data_ptr(), code_dev.numel(),
                                          N, K, out.data_ptr(), None, C.c_void_p(stream)), "cm_expand_stim_coo")
    return out


class CsrLam:
    """Sparse posterior of one fit: CSR over (neuron, trial) of the entries `lam` can be non-zero on.  `toarray()`
    gives the dense (N, K) float64 array the reference returns (caviar.py:100)."""

    def __init__(self, val, col, ptr, K):
        self.val, self.col, self.ptr, self.K = val, col, ptr, int(K)
        self.shape = (len(ptr) - 1, int(K))

    def toarray(self):
        val, col, ptr = (np.asarray(x.cpu() if hasattr(x, "cpu") else x) for x in (self.val, self.col, self.ptr))
        N = len(ptr) - 1
        out = np.zeros((N, self.K))
        nnz = int(ptr[-1])
        rows = np.repeat(np.arange(N), np.diff(ptr))
        out[rows, col[:nnz]] = val[:nnz]
        return out

    __array__ = lambda self, dtype=None, copy=None: self.toarray()


def caviar_batched(stim, powers, mu_prior, beta_prior, shape_prior, rate_prior, phi_prior, phi_cov_prior, psc=None,
                   y=None, ss=None, seeds=None, nnz_cap=None, want_lam=True, lam_csr=False, workspace=None, out=None,
                   cta_variant=0, **fit_options):
    """B fits on the current CUDA device.  All array arguments are CUDA tensors:
         stim (B,N,K) f32/f64 laser powers or uint8 power codes (c = powers[c-1], 0 = not targeted);
         psc (B,K,T) f32/f64 or (y, ss) (B,K) f64; priors (B,N[,2[,2]]) f64.
       `powers` is a host sequence (ascending distinct non-zero powers), `seeds` a host sequence of B ints.
       lam_csr=True adds the sparse posterior (lam_csr_val (B,nnz_cap), lam_csr_col, lam_csr_ptr (B,N+1)); `out` may
       carry preallocated output tensors from an earlier call with the same shapes (streaming reuses them).
       Returns a dict of CUDA tensors (mu, beta, lam, shape, rate, phi, phi_cov, z, status[, *_hist])."""
    import torch
    kw = _validate_options(fit_options)
    lib = _lib.load()
    dev = stim.device
    B, N, K = stim.shape
    if not stim.is_contiguous():
        raise ValueError("stim must be contiguous (B, N, K) with K fastest")
    iters = int(kw["iters"])
    f64 = dict(dtype=torch.float64, device=dev)
    powers = np.ascontiguousarray(powers, dtype=np.float64)
    if powers.size < 1 or powers.size > _lib.CM_CAVIAR_MAX_POWERS:
        raise RuntimeError("cm_caviar_fit supports 1..%d distinct stimulus powers, got %d"
                           % (_lib.CM_CAVIAR_MAX_POWERS, powers.size))
    if seeds is None:
        seeds = [int(kw["seed"])] * B
    seeds_arr = (C.c_uint64 * B)(*[int(s) & 0xFFFFFFFFFFFFFFFF for s in seeds])
    if nnz_cap is None:
        nnz_cap = int(torch.count_nonzero(stim.reshape(B, -1), dim=1).max().item())
    nnz_cap = max(int(nnz_cap), 1)
    hist = bool(kw["save_histories"])

    reuse = out if out is not None else {}

    def buf(name, shape, dtype=torch.float64):
        t = reuse.get(name)
        if t is not None and tuple(t.shape) == tuple(shape) and t.dtype == dtype and t.device == dev:
            return t
        return torch.empty(shape, dtype=dtype, device=dev)

    out = dict(mu=buf("mu", (B, N)), beta=buf("beta", (B, N)),
               lam=buf("lam", (B, N, K)) if want_lam else None,
               shape=buf("shape", (B,)), rate=buf("rate", (B,)),
               phi=buf("phi", (B, N, 2)), phi_cov=buf("phi_cov", (B, N, 2, 2)),
               z=buf("z", (B, K)), status=buf("status", (B,), torch.int32))
    if lam_csr:
        out.update(lam_csr_val=buf("lam_csr_val", (B, nnz_cap)), lam_csr_col=buf("lam_csr_col", (B, nnz_cap), torch.int32),
                   lam_csr_ptr=buf("lam_csr_ptr", (B, N + 1), torch.int32))
    if hist:
        out.update(mu_hist=torch.empty((B, iters, N), **f64), beta_hist=torch.empty((B, iters, N), **f64),
                   lam_hist=torch.empty((B, iters, N, K), **f64) if want_lam else None,
                   shape_hist=torch.empty((B, iters), **f64), rate_hist=torch.empty((B, iters), **f64),
                   phi_hist=torch.empty((B, iters, N, 2), **f64), phi_cov_hist=torch.empty((B, iters, N, 2, 2), **f64),
                   z_hist=torch.empty((B, iters, K), **f64))
    need = lib.cm_caviar_workspace_bytes(B, N, K, nnz_cap, iters if (hist and want_lam) else 0)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty((need,), dtype=torch.uint8, device=dev)

    a = _lib.CaviarArgs()
    a.B, a.N, a.K = B, N, K
    keep = []

    def ptr(t, dtype=None):
        if t is None:
            return None
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        t = t.contiguous()
        keep.append(t)
        return t.data_ptr()

    if psc is not None:
        if psc.shape[:2] != (B, K):
            raise ValueError("psc must be (B, K, T)")
        a.T = psc.shape[2]
        a.psc_dev, a.psc_dtype = ptr(psc), _dt(psc)
    else:
        a.T = 0
        a.y_dev, a.ss_dev = ptr(y, torch.float64), ptr(ss, torch.float64)
    a.stim_dev, a.stim_dtype = stim.data_ptr(), _dt(stim)
    a.n_powers = powers.size
    a.powers = powers.ctypes.data_as(C.POINTER(C.c_double))
    a.seeds = seeds_arr
    a.mu0_dev, a.beta0_dev = ptr(mu_prior, torch.float64), ptr(beta_prior, torch.float64)
    a.phi0_dev, a.phi_cov0_dev = ptr(phi_prior, torch.float64), ptr(phi_cov_prior, torch.float64)
    shape0 = (C.c_double * B)(*np.broadcast_to(np.asarray(shape_prior, dtype=np.float64), (B,)))
    rate0 = (C.c_double * B)(*np.broadcast_to(np.asarray(rate_prior, dtype=np.float64), (B,)))
    a.shape0, a.rate0 = shape0, rate0
    a.opt = _options(kw)
    for name in ("mu", "beta", "lam", "shape", "rate", "phi", "phi_cov", "z"):
        setattr(a, name + "_dev", out[name].data_ptr() if out[name] is not None else None)
    if hist:
        for name in ("mu", "beta", "lam", "shape", "rate", "phi", "phi_cov", "z"):
            t = out[name + "_hist"]
            setattr(a, name + "_hist_dev", t.data_ptr() if t is not None else None)
    if lam_csr:
        a.lam_csr_val_dev, a.lam_csr_col_dev = out["lam_csr_val"].data_ptr(), out["lam_csr_col"].data_ptr()
        a.lam_csr_ptr_dev = out["lam_csr_ptr"].data_ptr()
    a.cta_variant = int(cta_variant)
    a.nnz_cap = nnz_cap
    a.workspace_dev, a.workspace_bytes = workspace.data_ptr(), workspace.numel()
    a.status_dev = out["status"].data_ptr()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.cm_caviar_fit(C.byref(a), C.c_void_p(stream)), "cm_caviar_fit")
    out["_workspace"] = workspace
    out["_keep"] = keep
    out["launches"] = lib.cm_last_launch_count()
    return out


def check_status(out):
    st = out["status"].cpu().numpy()
    if np.any(st != 0):
        b = int(np.nonzero(st)[0][0])
        msg = {1: "stimulus matrix holds a value that is negative, NaN or not among `powers`",
               5: "non-zeros of the stimulus matrix exceed nnz_cap",
               9: "a helper CTA of the fit never answered; the results are not valid",
               10: "the two chain teams of the sweep lost each other; the results are not valid"}.get(int(st[b]), "device error")
        raise RuntimeError("cm_caviar_fit: fit %d failed with code %d (%s)" % (b, int(st[b]), msg))


def caviar(y_psc, I, mu_prior, beta_prior, shape_prior, rate_prior, phi_prior, phi_cov_prior, device=None,
           **fit_options):
    """Drop-in for circuitmap.optimise.caviar (caviar.py:20-100): NumPy in, the 17-tuple of NumPy arrays out."""
    _validate_options(fit_options)
    torch = _lib.require_cuda()
    print("Running coordinate-ascent variational inference and isotonic regularisation (CAVIaR) algorithm.")
    dev = torch.device("cuda" if device is None else device)
    y_dev = ss_dev = None
    with torch.cuda.device(dev):
        # ---- stimulus design: NumPy (the reference's surface) or a CUDA tensor that is already on the device ----
        if isinstance(I, torch.Tensor) and I.is_cuda:
            stim_dev = I.contiguous()
            if stim_dev.dtype not in (torch.float32, torch.float64):
                stim_dev = stim_dev.double()
            N, K = stim_dev.shape
            nnz, values = scan_stim(stim_dev)                     # np.unique(I)[1:] in one device pass (csrc/stim.cu)
            powers = powers_like_reference(nnz, values, N * K)
        else:
            I = np.asarray(I)
            if I.dtype not in (np.float32, np.float64):
                I = I.astype(float)
            I = np.ascontiguousarray(I)                  # simulate() returns a non-contiguous (N, K) view
            N, K = I.shape
            codes, powers, nnz = pack_stim_host(I)       # threaded C helper: powers + uint8 codes (N*K bytes to upload)
            stim_dev = codes.to(dev, non_blocking=True)
        # ---- traces: NumPy, or a CUDA tensor (e.g. straight from NeuralDemixer, carrying its y / ss statistics) ----
        if isinstance(y_psc, torch.Tensor) and y_psc.is_cuda:
            psc_dev = y_psc.contiguous()
            if psc_dev.dtype not in (torch.float32, torch.float64):
                psc_dev = psc_dev.double()
            y_dev, ss_dev = getattr(y_psc, "cm_y", None), getattr(y_psc, "cm_ss", None)
        else:
            y_psc = np.ascontiguousarray(y_psc)
            if y_psc.dtype not in (np.float32, np.float64):
                y_psc = y_psc.astype(float)
            psc_dev = _staging.to_device(y_psc, dev)
    if powers.size < 1:
        raise RuntimeError("cm_caviar_fit supports 1..%d distinct stimulus powers, got 0" % _lib.CM_CAVIAR_MAX_POWERS)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(dev)
    seed = int(fit_options.get("seed", _DEFAULTS["seed"]))
    opts = {k: v for k, v in fit_options.items() if k != "seed"}
    src = dict(y=y_dev[None], ss=ss_dev[None]) if (y_dev is not None and ss_dev is not None) else dict(psc=psc_dev[None])
    out = caviar_batched(stim_dev[None], powers, t(mu_prior)[None], t(beta_prior)[None],
                         float(shape_prior), float(rate_prior), t(phi_prior)[None], t(phi_cov_prior)[None],
                         seeds=[seed], nnz_cap=nnz, seed=seed, **src, **opts)
    check_status(out)
    def g(k):
        with torch.cuda.device(dev):
            return _staging.to_numpy(out[k][0])
    mu, beta, lam, phi, phi_cov, z = g("mu"), g("beta"), g("lam"), g("phi"), g("phi_cov"), g("z")
    shape, rate = np.float64(out["shape"][0].item()), np.float64(out["rate"][0].item())
    if out.get("mu_hist") is not None:
        iters = out["mu_hist"].shape[1]
        hist = [g("mu_hist"), g("beta_hist"), g("lam_hist"),
                np.repeat(g("shape_hist")[:, None], K, axis=1), np.repeat(g("rate_hist")[:, None], K, axis=1),
                g("phi_hist"), g("phi_cov_hist"), g("z_hist")]                 # caviar.py:57-64,92
    else:
        hist = [None] * 8
    return (mu, beta, lam, shape, rate, phi, phi_cov, z, None, *hist)
