"""Debug helper (not a test): accuracy + timing + phase breakdown of the multi-trace fp16 tensor-core demixer."""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from circuitmap_b200 import NeuralDemixer, _lib
from oracle import nwd as onwd
from oracle.make_golden import synth_traces
lib = _lib.load()
W = "tests/golden/nwd_ie_ChroME2f_weights.npz"
sd = dict(np.load(W)); folded = onwd.fold_bn(sd)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 203
traces = synth_traces(n, seed=21)
tmax = traces.max(1)[:, None]
ref = onwd.demix_np(traces.copy(), folded, monotone_start=900) / tmax
dem = NeuralDemixer(path=W, precision="fp16")
out = dem(traces.copy(), monotone_filter_start=900, verbose=False) / tmax
err = np.abs(out - ref)
print("fp16 mt path: nan rows %d; max-abs %.3e, pooled rel-L2 %.3e" % (np.isnan(out).any(1).sum(), np.nanmax(err),
      np.sqrt(np.nansum(err ** 2)) / np.sqrt((ref ** 2).sum())))
print("per-trace max err (first 12):", np.array2string(err.max(1)[:12], precision=4))
worst = int(np.nanargmax(err.max(1)))
print("worst trace %d at t=%d: out %.5f ref %.5f" % (worst, int(err[worst].argmax()), out[worst, err[worst].argmax()], ref[worst, err[worst].argmax()]))
dem.set_precision("tf32")
out1 = dem(traces.copy(), monotone_filter_start=900, verbose=False) / tmax
print("tf32 path: max-abs %.3e pooled rel-L2 %.3e" % (np.abs(out1 - ref).max(), np.sqrt(((out1 - ref) ** 2).sum()) / np.sqrt((ref ** 2).sum())))
dem.set_precision("fp16")
outf = dem(traces.copy(), verbose=False)
reff = onwd.demix_np(traces.copy(), folded)
print("with filter: max-abs(unit) %.3e; monotone %s" % (np.abs(outf - reff).max() / 1.0 if False else (np.abs(outf - reff) / tmax).max(), bool(np.all(np.diff(outf[:, 499:], axis=1) <= 0))))

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
x = torch.rand(K, 900, device="cuda")
o = torch.empty_like(x)
for prec in ("tf32", "fp16"):
    dem.set_precision(prec)
    for _ in range(3): dem.forward_device(x, out=o)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        dem.forward_device(x, out=o); ms.append(lib.cm_last_main_kernel_ms())
    m = float(np.mean(ms))
    print("%s: %.3f ms / %d traces = %.3f M traces/s, %.1f TFLOP/s" % (prec, m, K, K / m / 1e3, 2 * 8435200 * K / m / 1e9))
lib.cm_nwd_mt_debug_cycles(None, 0, 1)
dem.forward_device(x, out=o); torch.cuda.synchronize()
buf = (C.c_longlong * 32)()
lib.cm_nwd_mt_debug_cycles(buf, 32, 0)
names = ["input+P1+zeroC", "d1 mma", "d1 epi", "pool2", "d2 mma", "d2 epi", "pool3+d3+epi", "pool4+d4+epi", "u1+epi+interp1",
         "u2+epi+interp2", "u3 mma", "u3 epi", "interp3", "u4 mma", "u4 epi", "interp4", "fin mma", "fin epi", "output"]
npass = ((K + 3) // 4 + 147) // 148
tot = sum(buf[:19])
for i, nm in enumerate(names):
    print("  %-18s %8.0f cycles/pass %5.1f%%" % (nm, buf[i] / npass, 100.0 * buf[i] / max(tot, 1)))
print("  total %.0f cycles/pass = %.0f cycles/trace" % (tot / npass, tot / npass / 4))
print("  issuer: wait for weights per load (d1 d2 d3 d4 u1 u2 u3 u4+fin):", " ".join("%.0f" % (buf[19 + i] / npass) for i in range(8)))
print("  issuer: wait for A %.0f, wait for MMA completion %.0f cycles/pass" % (buf[28] / npass, buf[29] / npass))
