"""Debug helper (not a test): time cm_caviar_fit on synthetic maps."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import simulate as osim
from circuitmap_b200 import optimise

N, K, H, B, iters = [int(x) for x in sys.argv[1:6]]
t = time.time()
sim = osim.simulate_fast(N=N, K=K, H=H, seed=0)
print("simulate_fast %.1fs" % (time.time() - t))
stim = torch.from_numpy(sim["stim_matrix"]).cuda()[None].repeat(B, 1, 1).contiguous()
psc = torch.from_numpy(sim["psc"]).float().cuda()[None].repeat(B, 1, 1).contiguous()
ones = lambda *s: torch.ones(*s, dtype=torch.float64, device="cuda")
mu0 = torch.zeros(B, N, dtype=torch.float64, device="cuda"); beta0 = 10 * ones(B, N)
phi0 = torch.stack([0.1 * ones(B, N), 5 * ones(B, N)], -1).contiguous()
cov0 = torch.zeros(B, N, 2, 2, dtype=torch.float64, device="cuda"); cov0[..., 0, 0] = 0.1; cov0[..., 1, 1] = 1.0
powers = np.unique(sim["stim_matrix"])[1:]
ws = None
for rep in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    out = optimise.caviar_batched(stim, powers, mu0, beta0, 1.0, 0.1, phi0, cov0, psc=psc, seeds=list(range(1, B + 1)),
                                  nnz_cap=int(np.count_nonzero(sim["stim_matrix"])), want_lam=(B == 1), workspace=ws,
                                  iters=iters, msrmp=0.4)
    ws = out["_workspace"]
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("N=%d K=%d B=%d iters=%d: %.2f ms -> %.2f fits/s, %.1f iters/s; status %s; connected %d (truth %d)" % (
        N, K, B, iters, ms, B / ms * 1e3, B * iters / ms * 1e3, out["status"].sum().item(),
        int((out["mu"][0] != 0).sum()), int((sim["weights"] != 0).sum())))
tw = set(np.nonzero(sim["weights"])[0]); gw = set(np.nonzero(out["mu"][0].cpu().numpy())[0])
print("TP %d FP %d FN %d" % (len(tw & gw), len(gw - tw), len(tw - gw)))

# ---- phase breakdown of fit 0 ----
import ctypes as C
from circuitmap_b200 import _lib
lib = _lib.load()
import os
lib.cm_caviar_debug_phase_cycles(None, 0, int(os.environ.get('CM_DBG', '1')))
out = optimise.caviar_batched(stim, powers, mu0, beta0, 1.0, 0.1, phi0, cov0, psc=psc, seeds=list(range(1, B + 1)),
                              nnz_cap=int(np.count_nonzero(sim["stim_matrix"])), want_lam=False, workspace=ws, iters=iters, msrmp=0.4)
torch.cuda.synchronize()
buf = (C.c_longlong * 32)()
lib.cm_caviar_debug_phase_cycles(buf, 32, 0)
names = ["a2.compact+rows", "a2.gram", "a2.gemm1", "a2.S+chol", "a2.gemm2", "a2.Xupd", "a2.mu/beta", "a3.order", "a3.mc",
         "a3.pred+cst", "a3.sweep", "a6", "a7.newton", "a8", "hist+recon", "init"]
tot = sum(buf[:16])
for i, nm in enumerate(names):
    print("  %-16s %10.3f ms  %5.1f%%" % (nm, buf[i] / 1.9e6, 100.0 * buf[i] / max(tot, 1)))
print("  total %.2f ms (at 1.9 GHz)" % (tot / 1.9e6))
print("  chain steps %d, mean row len %.1f; cycles/step: pass1 %.0f reduce %.0f decide %.0f pass2 %.0f" % (
    buf[19], buf[24] / max(buf[19], 1), buf[20] / max(buf[19], 1), buf[21] / max(buf[19], 1), buf[22] / max(buf[19], 1), buf[23] / max(buf[19], 1)))
for i, nm in [(16, "sweep.chain warp"), (17, "sweep.rng warp"), (18, "sweep.inactive warp1")]:
    print("  %-22s %10.3f ms" % (nm, buf[i] / 1.9e6))
print("  newton: warp trips (one objective + gradient evaluation per row in flight) %d, evaluations the rows needed %d (%.2f per trip of 8 rows)" % (
    buf[25], buf[26], buf[26] / max(buf[25], 1)))
