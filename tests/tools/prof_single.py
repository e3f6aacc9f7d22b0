"""Profiling workload (not a test): ONE C3-shaped CAVIaR fit alone on the GPU (device-generated map), for ncu captures."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from circuitmap_b200 import optimise
from circuitmap_b200.simulation import simulate_batch

N, K = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 10000)
r = simulate_batch([1], N=N, trials=K, H=10, connection_prob=0.1)
f64 = dict(dtype=torch.float64, device="cuda")
cov = torch.zeros(1, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
phi = torch.stack([0.1 * torch.ones(1, N, **f64), 5 * torch.ones(1, N, **f64)], -1).contiguous()
ws = None
for rep in range(2):
    out = optimise.caviar_batched(r["codes"], r["powers"], torch.zeros(1, N, **f64), 10 * torch.ones(1, N, **f64), 1.0, 0.1, phi, cov,
                                  psc=r["psc"], seeds=[1], nnz_cap=K * 10, want_lam=False, lam_csr=True, workspace=ws, iters=50, msrmp=0.4)
    ws = out["_workspace"]
torch.cuda.synchronize()
print("caviar ok", int(out["status"].sum()), int((out["mu"][0] != 0).sum()), "of", int((r["weights"][0] != 0).sum()))
