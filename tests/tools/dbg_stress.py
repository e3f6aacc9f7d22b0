"""Debug helper (not a test): repeat the same batched fit and check run-to-run bitwise equality of every fit."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import simulate as osim
from circuitmap_b200 import optimise
N, K, H, B, iters, reps = [int(x) for x in sys.argv[1:7]]
sim = osim.simulate_fast(N=N, K=K, H=H, seed=0)
stim = torch.from_numpy(sim["stim_matrix"]).cuda()[None].repeat(B, 1, 1).contiguous()
psc = torch.from_numpy(sim["psc"]).float().cuda()[None].repeat(B, 1, 1).contiguous()
f64 = dict(dtype=torch.float64, device="cuda")
cov = torch.zeros(B, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
phi = torch.stack([0.1 * torch.ones(B, N, **f64), 5 * torch.ones(B, N, **f64)], -1).contiguous()
pri = (torch.zeros(B, N, **f64), 10 * torch.ones(B, N, **f64), 1.0, 0.1, phi, cov)
powers = np.unique(sim["stim_matrix"])[1:]
ws = None; ref = None
for r in range(reps):
    out = optimise.caviar_batched(stim, powers, *pri, psc=psc, seeds=[1] * B, nnz_cap=int(np.count_nonzero(sim["stim_matrix"])),
                                  want_lam=False, workspace=ws, iters=iters, msrmp=0.4)
    ws = out["_workspace"]
    torch.cuda.synchronize()
    mu = out["mu"].clone()
    same_within = bool((mu == mu[0:1]).all())          # all fits identical inputs+seed -> identical outputs
    if ref is None: ref = mu
    print("rep %d: all fits equal to fit 0: %s; equal to rep 0: %s; connected fit0 %d; distinct connected counts %s" % (
        r, same_within, bool((mu == ref).all()), int((mu[0] != 0).sum()), sorted(set((mu != 0).sum(1).tolist()))))
