"""Drop-in `NeuralDemixer` (reference circuitmap/neural_waveform_demixing.py:17-54) on the B200 kernel.

Same constructor and call signature, same float64 (K, 900) result; the U-Net forward, the per-trace
normalisation and the monotone decay filter all run inside `cm_nwd_forward` (csrc/nwd.cu).  Training
(`train`, `generate_training_data`, nwd.py:56-180) is out of scope for the hot path.
"""
import ctypes as C
import time

import numpy as np

from . import _lib

# state_dict key order expected by cm_nwd_create (nwd.py:259-269)
_BLOCKS = [("dblock1", "conv"), ("dblock2", "conv"), ("dblock3", "conv"), ("dblock4", "conv"),
           ("ublock1", "deconv"), ("ublock2", "deconv"), ("ublock3", "deconv"), ("ublock4", "deconv"),
           ("conv", "conv")]
_SHAPES = [(16, 1, 32), (16, 16, 32), (32, 16, 16), (32, 32, 16), (32, 16, 16), (48, 16, 16), (32, 16, 32),
           (32, 4, 32), (1, 4, 256)]


def state_dict_keys():
    keys = []
    for blk, cname in _BLOCKS:
        keys += [f"{blk}.{cname}.weight", f"{blk}.{cname}.bias", f"{blk}.bn.weight", f"{blk}.bn.bias",
                 f"{blk}.bn.running_mean", f"{blk}.bn.running_var"]
    return keys


def load_weights(path):
    """Read a Lightning .ckpt (state_dict) or an .npz of the same tensors -> {key: float32 ndarray}."""
    if str(path).endswith(".npz"):
        sd = dict(np.load(path))
    else:
        import torch
        ck = torch.load(path, map_location="cpu", weights_only=True)
        sd = ck["state_dict"] if "state_dict" in ck else ck
        sd = {k: v.numpy() for k, v in sd.items()}
    return {k: np.ascontiguousarray(sd[k], dtype=np.float32) for k in state_dict_keys()}


def random_weights(seed=0):
    """Random-init weights of the NWDUNet architecture (PyTorch default conv init, BN at its eval defaults)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for (blk, cname), shp in zip(_BLOCKS, _SHAPES):
        fan_in = (shp[1] if cname == "conv" else shp[0]) * shp[2]
        cout = shp[0] if cname == "conv" else shp[1]
        bound = 1.0 / np.sqrt(fan_in)
        sd[f"{blk}.{cname}.weight"] = rng.uniform(-bound, bound, shp).astype(np.float32)
        sd[f"{blk}.{cname}.bias"] = rng.uniform(-bound, bound, cout).astype(np.float32)
        sd[f"{blk}.bn.weight"] = np.ones(cout, np.float32)
        sd[f"{blk}.bn.bias"] = np.zeros(cout, np.float32)
        sd[f"{blk}.bn.running_mean"] = np.zeros(cout, np.float32)
        sd[f"{blk}.bn.running_var"] = np.ones(cout, np.float32)
    return sd


class NeuralDemixer:
    def __init__(self, path=None, eval_mode=True, device=None, precision="fp32", state_dict=None):
        """precision (extension of the reference signature): 'fp32' = fp32 CUDA-core convolutions (default, the
        reference's arithmetic); 'tf32' = tcgen05 tensor-core path, one trace per CTA (TF32 operands, fp32
        accumulate); 'fp16' = the fast tcgen05 path: all nine convolutions as widened implicit GEMMs with several
        traces per M tile (fp16 operands = TF32's 11-bit significand, fp32 accumulate).  Both tensor-core modes:
        max-abs error <= 2e-2 on unit-normalised traces."""
        torch = _lib.require_cuda()
        self.device = torch.device("cuda" if device is None else device)
        if self.device.type != "cuda":
            raise RuntimeError("circuitmap_b200.NeuralDemixer runs on CUDA only (no CPU fallback); got device=%r"
                               % (device,))
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if not eval_mode:
            raise NotImplementedError("training-mode BatchNorm (eval_mode=False) is outside the inference hot path")
        if state_dict is not None:        # extension: weights given directly ({state_dict key: float array})
            self.weights = {k: np.ascontiguousarray(state_dict[k], dtype=np.float32) for k in state_dict_keys()}
        else:
            self.weights = load_weights(path) if path is not None else random_weights()
        self._lib = _lib.load()
        ptrs = (C.c_void_p * _lib.CM_NWD_NUM_TENSORS)(
            *[self.weights[k].ctypes.data_as(C.c_void_p) for k in state_dict_keys()])
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cm_nwd_create(ptrs, _lib.CM_NWD_NUM_TENSORS, C.byref(h)), "cm_nwd_create")
        self._h = h
        self.set_precision(precision)

    def set_precision(self, precision):
        mode = {"fp32": 0, "tf32": 1, "fp16": 2}.get(precision)
        if mode is None:
            raise ValueError("precision must be 'fp32', 'tf32' or 'fp16'")
        _lib.check(self._lib.cm_nwd_set_precision(self._h, mode), "cm_nwd_set_precision")
        self.precision = precision

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.cm_nwd_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- device-resident entry point (used by Model.fit hand-off and bench.py) -------------------------------
    def forward_device(self, traces, monotone_filter_start=500, out=None, out_dtype=None, stats=False):
        """traces: CUDA tensor (K, 900) float32/float64 -> demixed CUDA tensor.  With stats=True also returns
        (y, ss): per-trace trapz and sum of squares (the CAVIaR prologue, caviar.py:28-30)."""
        import torch
        if not (traces.is_cuda and traces.dim() == 2 and traces.is_contiguous()):
            raise ValueError("traces must be a contiguous 2-D CUDA tensor")
        K, T = traces.shape
        in_dt = {torch.float32: _lib.CM_F32, torch.float64: _lib.CM_F64}[traces.dtype]
        if out is None:
            out = torch.empty((K, T), dtype=out_dtype or traces.dtype, device=traces.device)
        out_dt = {torch.float32: _lib.CM_F32, torch.float64: _lib.CM_F64}[out.dtype]
        y = ss = None
        if stats:
            y = torch.empty(K, dtype=torch.float64, device=traces.device)
            ss = torch.empty(K, dtype=torch.float64, device=traces.device)
        with torch.cuda.device(traces.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = self._lib.cm_nwd_forward(self._h, traces.data_ptr(), in_dt, out.data_ptr(), out_dt, K, T,
                                          int(monotone_filter_start), y.data_ptr() if stats else None,
                                          ss.data_ptr() if stats else None, C.c_void_p(stream))
        _lib.check(rc, "cm_nwd_forward")
        return (out, y, ss) if stats else out

    # ---- reference call surface (nwd.py:36-54) -----------------------------------------------------------------
    def __call__(self, traces, monotone_filter_start=500, monotone_filter_inplace=True, verbose=True):
        """Run demixer over PSC trace batch and apply monotone decay filter."""
        import torch
        if verbose:
            print("Demixing PSC traces... ", end="")
        t1 = time.time()
        if isinstance(traces, torch.Tensor):
            # device tensor in, device tensor out: the demixed traces stay on the GPU and carry the CAVIaR prologue
            # statistics (y = trapz, sum of squares; caviar.py:28-30) so that Model.fit(dem, ...) neither re-reads the
            # K x 900 array nor moves it through the host
            x = traces.to(self.device).contiguous()
            if x.dtype not in (torch.float32, torch.float64):
                x = x.double()
            dem, y, ss = self.forward_device(x, monotone_filter_start, stats=True)
            dem.cm_y, dem.cm_ss = y, ss
        else:
            arr = np.ascontiguousarray(traces)
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64)
            squeeze = arr.ndim == 1
            arr2 = arr.reshape(-1, arr.shape[-1])
            if arr2.shape[0] > self._PIPE_CHUNK and arr2.shape[1] == _lib.CM_NWD_T:
                dem = self._call_pipelined(arr2, monotone_filter_start)
            else:
                x = torch.from_numpy(arr2).to(self.device)
                dem = self.forward_device(x, monotone_filter_start, out_dtype=torch.float64).cpu().numpy()
            if squeeze:
                dem = dem[0]
        t2 = time.time()
        if verbose:
            print("complete (elapsed time %.2fs, device=%s)." % (t2 - t1, self.device))
        return dem

    # ---- large NumPy batches: chunked through two pinned staging buffers ------------------------------------------
    _PIPE_CHUNK = 4096

    def _call_pipelined(self, arr, monotone_filter_start):
        """NumPy (K, 900) -> float64 NumPy, K > _PIPE_CHUNK: the host copies into / out of pinned staging, the
        host->device copy, the kernel and the device->host copy of consecutive chunks overlap (results identical to the
        one-shot path: traces are independent)."""
        import torch
        K, T = arr.shape
        C = self._PIPE_CHUNK
        st = getattr(self, "_pipe", None)
        if st is None or st["dtype"] != arr.dtype:
            tdt = torch.float32 if arr.dtype == np.float32 else torch.float64
            st = dict(dtype=arr.dtype,
                      pin_in=[torch.empty((C, T), dtype=tdt).pin_memory() for _ in range(2)],
                      pin_out=[torch.empty((C, T), dtype=torch.float64).pin_memory() for _ in range(2)],
                      dev_in=[torch.empty((C, T), dtype=tdt, device=self.device) for _ in range(2)],
                      dev_out=[torch.empty((C, T), dtype=torch.float64, device=self.device) for _ in range(2)],
                      streams=[torch.cuda.Stream(device=self.device) for _ in range(3)])
            self._pipe = st
        s_in, s_k, s_out = st["streams"]
        cur = torch.cuda.current_stream(self.device)
        for s_ in (s_in, s_k, s_out):
            s_.wait_stream(cur)
        out = np.empty((K, T), dtype=np.float64)
        src = torch.from_numpy(arr)
        dst = torch.from_numpy(out)
        ev_h2d, ev_k, ev_d2h = [None, None], [None, None], [None, None]
        bounds = [(lo, min(lo + C, K)) for lo in range(0, K, C)]
        for i, (lo, hi) in enumerate(bounds):
            b, n = i & 1, hi - lo
            if ev_h2d[b] is not None:
                ev_h2d[b].synchronize()                       # staging buffer b is free again
            st["pin_in"][b][:n].copy_(src[lo:hi])
            with torch.cuda.stream(s_in):
                if ev_k[b] is not None:
                    s_in.wait_event(ev_k[b])                  # the kernel that read dev_in[b] two chunks ago is done
                st["dev_in"][b][:n].copy_(st["pin_in"][b][:n], non_blocking=True)
                ev_h2d[b] = s_in.record_event()
            with torch.cuda.stream(s_k):
                s_k.wait_event(ev_h2d[b])
                if ev_d2h[b] is not None:
                    s_k.wait_event(ev_d2h[b])                 # dev_out[b] has been copied out
                self.forward_device(st["dev_in"][b][:n], monotone_filter_start, out=st["dev_out"][b][:n])
                ev_k[b] = s_k.record_event()
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_k[b])
                st["pin_out"][b][:n].copy_(st["dev_out"][b][:n], non_blocking=True)
                ev_d2h[b] = s_out.record_event()
            if i >= 1:                                        # drain the previous chunk while this one is in flight
                plo, phi = bounds[i - 1]
                ev_d2h[1 - b].synchronize()
                dst[plo:phi].copy_(st["pin_out"][1 - b][:phi - plo])
        plo, phi = bounds[-1]
        lb = (len(bounds) - 1) & 1
        ev_d2h[lb].synchronize()
        dst[plo:phi].copy_(st["pin_out"][lb][:phi - plo])
        cur.wait_stream(s_out)
        return out

    def train(self, *a, **k):
        raise NotImplementedError("demixer training (nwd.py:56-94) is outside the B200 inference hot path")

    def generate_training_data(self, *a, **k):
        raise NotImplementedError("training-data synthesis (nwd.py:96-163) is outside the B200 inference hot path")
