"""compute-sanitizer workload (not a test): small fits through every a2 path (small in-smem solve, tile solve with helper CTAs,
bordered recursion of the 8-warp variant), the device generator, the sparse design expansion and the demixer hand-off."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from circuitmap_b200 import NeuralDemixer, optimise
from circuitmap_b200.simulation import simulate_batch

def pri(B, N):
    f64 = dict(dtype=torch.float64, device="cuda")
    cov = torch.zeros(B, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
    phi = torch.stack([0.1 * torch.ones(B, N, **f64), 5 * torch.ones(B, N, **f64)], -1).contiguous()
    return torch.zeros(B, N, **f64), 10 * torch.ones(B, N, **f64), 1.0, 0.1, phi, cov

for (B, N, K, H, cta) in [(1, 40, 300, 4, 0), (1, 301, 2500, 7, 0), (3, 270, 2000, 6, 256), (2, 301, 2500, 7, 512)]:
    r = simulate_batch(list(range(1, B + 1)), N=N, trials=K, H=H, connection_prob=0.1)
    out = optimise.caviar_batched(r["codes"], r["powers"], *pri(B, N), psc=r["psc"], seeds=list(range(B)), nnz_cap=K * H,
                                  want_lam=(B == 1), lam_csr=True, cta_variant=cta, iters=5, msrmp=0.4)
    torch.cuda.synchronize()
    print("fit", B, N, K, cta, "status", int(out["status"].sum()), "connected", int((out["mu"][0] != 0).sum()))
    n, t, c = optimise.codes_to_coo(r["codes"][0])
    back = optimise.expand_coo(n.cuda(), t.cuda(), c.cuda(), N, K)
    assert torch.equal(back, r["codes"][0])
dem = NeuralDemixer(path="tests/golden/nwd_ie_ChroME2f_weights.npz", precision="fp16")
d = dem(r["psc"][0][:64].contiguous(), verbose=False)
torch.cuda.synchronize()
print("demix ok", tuple(d.shape), float(d.cm_y.abs().max()))
