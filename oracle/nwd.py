"""NWD U-Net forward restated (oracle; test infrastructure only).

Restates circuitmap/neural_waveform_demixing.py:
  NeuralDemixer.__call__          :36-54   (normalise by per-trace max, net, rescale, filter)
  DownsamplingBlock.forward       :204-217 (relu(bn(conv(avgpool3s2(x)))))
  UpsamplingBlock.forward         :219-238 (interp(relu(bn(deconv(x)))) [cat skip])
  ConvolutionBlock.forward        :240-252
  NWDUNet.__init__/forward        :254-287
  _monotone_decay_filter          :337-348

Two independent restatements:
  forward_np   explicit NumPy formulas in fp64 (BN folded; pooling / transposed conv / linear
               interpolation written out) -- pins the layer maths the CUDA kernel implements;
  TorchNWD     the same network from torch.nn.functional ops in fp32 -- the CPU baseline and the
               fp32 comparison point.
Both are checked against the UNMODIFIED reference module (imported through a pytorch_lightning
shim by oracle/make_golden.py) via the golden vectors in tests/golden/.  Parity: PINNED.
"""
import numpy as np

# (name, kind, c_in, c_out, k, dilation, stride, padding)   nwd.py:259-269
LAYERS = [
    ("dblock1", "down", 1, 16, 32, 2, 1, 0),
    ("dblock2", "down", 16, 16, 32, 1, 1, 0),
    ("dblock3", "down", 16, 32, 16, 1, 1, 0),
    ("dblock4", "down", 32, 32, 16, 1, 1, 0),
    ("ublock1", "up", 32, 16, 16, 1, 1, 0),
    ("ublock2", "up", 48, 16, 16, 1, 1, 0),
    ("ublock3", "up", 32, 16, 32, 1, 1, 0),
    ("ublock4", "up", 32, 4, 32, 1, 2, 0),
    ("conv", "final", 4, 1, 256, 2, 1, 255),
]
BN_EPS = 1e-5


def load_state_dict(path):
    """state_dict of a Lightning checkpoint as {name: float64/float32 ndarray} (nwd.py:27)."""
    import torch
    ck = torch.load(path, map_location="cpu", weights_only=True)
    sd = ck["state_dict"] if "state_dict" in ck else ck
    return {k: v.numpy() for k, v in sd.items()}


def fold_bn(sd, dtype=np.float64):
    """Fold eval-mode BatchNorm into each conv: W' = W*g/sqrt(rv+eps), b' = (b-rm)*g/sqrt(rv+eps)+beta."""
    out = {}
    for name, kind, cin, cout, k, dil, stride, pad in LAYERS:
        cname = "deconv" if kind == "up" else "conv"
        W = sd[f"{name}.{cname}.weight"].astype(np.float64)
        b = sd[f"{name}.{cname}.bias"].astype(np.float64)
        g = sd[f"{name}.bn.weight"].astype(np.float64)
        be = sd[f"{name}.bn.bias"].astype(np.float64)
        rm = sd[f"{name}.bn.running_mean"].astype(np.float64)
        rv = sd[f"{name}.bn.running_var"].astype(np.float64)
        s = g / np.sqrt(rv + BN_EPS)
        if kind == "up":                      # ConvTranspose1d weight is (c_in, c_out, k)
            Wf = W * s[None, :, None]
        else:
            Wf = W * s[:, None, None]
        out[name] = (Wf.astype(dtype), ((b - rm) * s + be).astype(dtype))
    return out


def avgpool3s2(x):
    L = x.shape[-1]
    Lo = (L - 3) // 2 + 1
    i = 2 * np.arange(Lo)
    return (x[..., i] + x[..., i + 1] + x[..., i + 2]) / 3.0


def conv1d(x, W, b, dil=1, pad=0):
    """x (B,Ci,L), W (Co,Ci,k) -> (B,Co,L+2pad-dil(k-1))."""
    if pad:
        x = np.pad(x, ((0, 0), (0, 0), (pad, pad)))
    B, Ci, L = x.shape
    Co, _, k = W.shape
    Lo = L - dil * (k - 1)
    out = np.zeros((B, Co, Lo), dtype=x.dtype)
    for j in range(k):
        out += np.einsum("oc,bcl->bol", W[:, :, j], x[:, :, j * dil:j * dil + Lo])
    return out + b[None, :, None]


def conv_transpose1d(x, W, b, stride=1):
    """x (B,Ci,L), W (Ci,Co,k): out[co, s*i+j] += x[ci,i]*W[ci,co,j]."""
    B, Ci, L = x.shape
    _, Co, k = W.shape
    Lo = (L - 1) * stride + k
    out = np.zeros((B, Co, Lo), dtype=x.dtype)
    for j in range(k):
        out[:, :, j:j + (L - 1) * stride + 1:stride] += np.einsum("co,bcl->bol", W[:, :, j], x)
    return out + b[None, :, None]


def interp_linear(x, size):
    """F.interpolate(mode='linear', align_corners=False)."""
    L = x.shape[-1]
    scale = L / size
    src = np.maximum((np.arange(size) + 0.5) * scale - 0.5, 0.0)
    i0 = np.floor(src).astype(int)
    i1 = np.minimum(i0 + 1, L - 1)
    w1 = src - i0
    return (1 - w1) * x[..., i0] + w1 * x[..., i1]


def forward_np(x, folded):
    """NWDUNet.forward (nwd.py:271-287) on x (B, T) -> (B, T); dtype follows x."""
    relu = lambda a: np.maximum(a, 0)
    x = x[:, None, :]
    T = x.shape[-1]
    enc = []
    h = x
    for name, dil in (("dblock1", 2), ("dblock2", 1), ("dblock3", 1), ("dblock4", 1)):
        W, b = folded[name]
        h = relu(conv1d(avgpool3s2(h), W, b, dil=dil))
        enc.append(h)
    skips = [enc[2], enc[1], enc[0], None]
    for name, stride, skip in zip(("ublock1", "ublock2", "ublock3", "ublock4"), (1, 1, 1, 2), skips):
        W, b = folded[name]
        up = relu(conv_transpose1d(h, W, b, stride=stride))
        if skip is not None:
            h = np.concatenate([interp_linear(up, skip.shape[-1]), skip], axis=1)
        else:
            h = interp_linear(up, T)
    W, b = folded["conv"]
    out = relu(conv1d(h, W, b, dil=2, pad=255))
    return out[:, 0, :]


def monotone_decay_filter(arr, monotone_start=500):
    """nwd.py:337-348 (in place): running minimum from column monotone_start-1 onwards."""
    if monotone_start < arr.shape[1]:
        s = max(monotone_start - 1, 0)
        arr[:, s:] = np.minimum.accumulate(arr[:, s:], axis=1)
    return arr


def demix_np(traces, folded, monotone_start=500, dtype=np.float64):
    """NeuralDemixer.__call__ (nwd.py:36-54) around forward_np."""
    traces = np.asarray(traces, dtype=np.float64)
    tmax = np.max(traces, axis=1)[:, None]
    x = (traces / tmax).astype(dtype)
    f = {k: (w.astype(dtype), b.astype(dtype)) for k, (w, b) in folded.items()}
    dem = forward_np(x, f).astype(np.float64) * tmax
    return monotone_decay_filter(dem, monotone_start)


class TorchNWD:
    """fp32 torch.nn.functional restatement (eval-mode BN kept un-folded, as the reference runs it)."""

    def __init__(self, sd):
        import torch
        self.t = torch
        self.sd = {k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()}

    def _bn(self, x, name):
        F = self.t.nn.functional
        s = self.sd
        return F.batch_norm(x, s[f"{name}.bn.running_mean"], s[f"{name}.bn.running_var"],
                            s[f"{name}.bn.weight"], s[f"{name}.bn.bias"], False, 0.0, BN_EPS)

    def forward(self, x):
        t, s = self.t, self.sd
        F = t.nn.functional
        with t.no_grad():
            T = x.shape[-1]
            enc = []
            h = x
            for name, dil in (("dblock1", 2), ("dblock2", 1), ("dblock3", 1), ("dblock4", 1)):
                h = F.avg_pool1d(h, 3, 2)
                h = F.relu(self._bn(F.conv1d(h, s[f"{name}.conv.weight"], s[f"{name}.conv.bias"], dilation=dil), name))
                enc.append(h)
            skips = [enc[2], enc[1], enc[0], None]
            for name, stride, skip in zip(("ublock1", "ublock2", "ublock3", "ublock4"), (1, 1, 1, 2), skips):
                up = F.relu(self._bn(F.conv_transpose1d(h, s[f"{name}.deconv.weight"], s[f"{name}.deconv.bias"],
                                                        stride=stride), name))
                size = skip.shape[-1] if skip is not None else T
                up = F.interpolate(up, size=size, mode="linear", align_corners=False)
                h = t.cat([up, skip], dim=1) if skip is not None else up
            out = F.relu(self._bn(F.conv1d(h, s["conv.conv.weight"], s["conv.conv.bias"], padding=255, dilation=2),
                                  "conv"))
        return out

    def demix(self, traces, monotone_start=500):
        """NeuralDemixer.__call__ (nwd.py:36-54): float64 in, float64 out, whole K in one batch."""
        t = self.t
        tmax = np.max(traces, axis=1)[:, None]
        x = t.Tensor((traces / tmax).copy()[:, None, :])
        dem = self.forward(x).cpu().numpy().squeeze(1) * tmax
        return monotone_decay_filter(dem, monotone_start)
