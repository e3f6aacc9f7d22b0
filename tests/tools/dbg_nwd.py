"""Debug helper (not a test): time cm_nwd_forward in both precisions."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from circuitmap_b200 import NeuralDemixer, _lib
lib = _lib.load()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
x = torch.rand(K, 900, device="cuda")
for prec in ("fp32", "tf32"):
    d = NeuralDemixer(path="tests/golden/nwd_ie_ChroME2f_weights.npz", precision=prec)
    o = torch.empty_like(x)
    for _ in range(3): d.forward_device(x, out=o)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        d.forward_device(x, out=o); ms.append(lib.cm_last_main_kernel_ms())
    m = float(np.mean(ms))
    print("%s: %.3f ms / %d traces = %.3f M traces/s, %.1f TFLOP/s" % (prec, m, K, K / m / 1e3, 2 * 8435200 * K / m / 1e9))

import ctypes as C
lib.cm_nwd_debug_cycles(None, 0, 1)
d.forward_device(x, out=o); torch.cuda.synchronize()
buf = (C.c_longlong * 16)()
lib.cm_nwd_debug_cycles(buf, 16, 0)
names = ["load+pool1+d1", "pool2+d2+epi", "pool3+d3+epi", "pool4+d4+epi", "u1+epi", "interp1+u2+epi", "interp2+u3+epi", "interp3+u4+epi", "interp4", "final conv", "filter+store"]
ntr = (K + 147) // 148
tot = sum(buf[:11])
for i, nm in enumerate(names):
    print("  %-18s %8.0f cycles/trace %5.1f%%" % (nm, buf[i] / ntr, 100.0 * buf[i] / tot))
print("  total %.0f cycles/trace" % (tot / ntr))
