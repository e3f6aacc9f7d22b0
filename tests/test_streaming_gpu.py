"""GPU: the chunked, copy/compute-overlapped host entry points (circuitmap_b200/streaming.py) return exactly what the
one-shot calls return."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_demix_pinned_equals_one_shot():
    import torch
    from circuitmap_b200 import NeuralDemixer, streaming
    from oracle.make_golden import synth_traces
    dem = NeuralDemixer(path=os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz"), precision="fp16")
    traces = synth_traces(1001, seed=8)
    src = torch.from_numpy(traces).pin_memory()
    dst = torch.empty_like(src).pin_memory()
    din = torch.empty(src.shape, dtype=torch.float64, device="cuda")
    dout = torch.empty_like(din)
    for chunk in (1001, 256, 7):
        dst.zero_()
        streaming.demix_pinned(dem, src, dst, din, dout, chunk=chunk)
        assert np.array_equal(dst.numpy(), dem(traces.copy(), verbose=False))


def test_fit_pinned_equals_batched_call():
    import torch
    from circuitmap_b200 import optimise, streaming
    from oracle import simulate as osim
    B, N, K = 5, 24, 200
    sims = [osim.simulate_fast(N=N, K=K, H=4, seed=s) for s in range(B)]
    f64 = dict(dtype=torch.float64, device="cuda")
    hs = [torch.from_numpy(np.ascontiguousarray(s["stim_matrix"])).pin_memory() for s in sims]
    hp = [torch.from_numpy(np.ascontiguousarray(s["psc"])).pin_memory() for s in sims]
    stim = torch.stack([h.cuda() for h in hs]).contiguous()
    psc = torch.stack([h.cuda() for h in hp]).contiguous()
    cov = torch.zeros(B, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
    phi = torch.stack([0.1 * torch.ones(B, N, **f64), 5 * torch.ones(B, N, **f64)], -1).contiguous()
    pri = (torch.zeros(B, N, **f64), 10 * torch.ones(B, N, **f64), 1.0, 0.1, phi, cov)
    powers = np.unique(sims[0]["stim_matrix"])[1:]
    seeds = [3, 1, 4, 1, 5]
    ref = optimise.caviar_batched(stim, powers, *pri, psc=psc, seeds=seeds, iters=12, msrmp=0.4)
    host_out = {k: torch.empty(ref[k].shape, dtype=ref[k].dtype).pin_memory()
                for k in ("mu", "beta", "shape", "rate", "phi", "phi_cov", "z", "lam")}
    dstim, dpsc = torch.empty_like(stim), torch.empty_like(psc)
    status = streaming.fit_pinned(hs, hp, dstim, dpsc, powers, pri, seeds, host_out, chunk=2, iters=12, msrmp=0.4)
    assert int(status.sum().item()) == 0 and status.numel() == B
    for k, v in host_out.items():
        assert np.array_equal(v.numpy(), ref[k].cpu().numpy(), equal_nan=True), k
