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
            dem = self.forward_device(traces.to(self.device).contiguous(), monotone_filter_start)
        else:
            arr = np.ascontiguousarray(traces)
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64)
            squeeze = arr.ndim == 1
            x = torch.from_numpy(arr.reshape(-1, arr.shape[-1])).to(self.device)
            dem = self.forward_device(x, monotone_filter_start, out_dtype=torch.float64).cpu().numpy()
            if squeeze:
                dem = dem[0]
        t2 = time.time()
        if verbose:
            print("complete (elapsed time %.2fs, device=%s)." % (t2 - t1, self.device))
        return dem

    def train(self, *a, **k):
        raise NotImplementedError("demixer training (nwd.py:56-94) is outside the B200 inference hot path")

    def generate_training_data(self, *a, **k):
        raise NotImplementedError("training-data synthesis (nwd.py:96-163) is outside the B200 inference hot path")
