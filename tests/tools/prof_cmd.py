"""Profiling workload (not a test): one C3-shaped CAVIaR batch and one C2 NWD forward, for ncu captures."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from bench import synth_map, synth_traces
from circuitmap_b200 import NeuralDemixer, optimise

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N, K = 1000, 10000
stim_h, psc_h, _ = synth_map(N, K, 10, seed=0)
f64 = dict(dtype=torch.float64, device="cuda")
stim = torch.from_numpy(stim_h).cuda()[None].repeat(B, 1, 1).contiguous()
psc = torch.from_numpy(psc_h).cuda()[None].repeat(B, 1, 1).contiguous()
cov = torch.zeros(B, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
phi = torch.stack([0.1 * torch.ones(B, N, **f64), 5 * torch.ones(B, N, **f64)], -1).contiguous()
out = optimise.caviar_batched(stim, np.array([45., 55., 65.]), torch.zeros(B, N, **f64), 10 * torch.ones(B, N, **f64), 1.0,
                              0.1, phi, cov, psc=psc, seeds=list(range(B)), iters=50, msrmp=0.4)
torch.cuda.synchronize()
print("caviar ok", int(out["status"].sum()), int((out["mu"][0] != 0).sum()))
dem = NeuralDemixer(path="tests/golden/nwd_ie_ChroME2f_weights.npz")
x = torch.from_numpy(synth_traces(20000, 0)).float().cuda()
o = dem.forward_device(x)
torch.cuda.synchronize()
print("nwd ok", float(o.abs().max()))
