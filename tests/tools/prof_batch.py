"""Profiling workload (not a test): B distinct C3-shaped maps (device generator) through ONE cm_caviar_fit launch, twice."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from circuitmap_b200 import optimise
from circuitmap_b200.simulation import simulate_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
N, K = 1000, 10000
codes = torch.empty((B, N, K), dtype=torch.uint8, device="cuda")
psc = torch.empty((B, K, 900), dtype=torch.float32, device="cuda")
ws = None
for lo in range(0, B, 32):
    hi = min(lo + 32, B)
    r = simulate_batch(list(range(1 + lo, 1 + hi)), workspace=ws, N=N, trials=K, H=10, connection_prob=0.1, out=dict(codes=codes[lo:hi], psc=psc[lo:hi]))
    ws = r["_workspace"]
f64 = dict(dtype=torch.float64, device="cuda")
cov = torch.zeros(B, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
phi = torch.stack([0.1 * torch.ones(B, N, **f64), 5 * torch.ones(B, N, **f64)], -1).contiguous()
w = None
for rep in range(2):
    out = optimise.caviar_batched(codes, r["powers"], torch.zeros(B, N, **f64), 10 * torch.ones(B, N, **f64), 1.0, 0.1, phi, cov,
                                  psc=psc, seeds=list(range(1, B + 1)), nnz_cap=K * 10, want_lam=False, lam_csr=True, workspace=w, iters=50, msrmp=0.4)
    w = out["_workspace"]
torch.cuda.synchronize()
print("caviar ok", int(out["status"].sum()), int((out["mu"][0] != 0).sum()))
