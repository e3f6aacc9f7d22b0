"""GPU parity: cm_caviar_fit (through Model.fit / optimise.caviar, i.e. the C ABI) vs the NumPy fp64 oracle.

Tolerance (BASELINE.json north_star): connected set and accept/reject pattern IDENTICAL; mu/beta/lam/phi/phi_cov/
shape/rate/z within 1e-4 relative.  The kernel computes in fp64 and lands around 1e-8.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-4
NAMES = ["mu", "beta", "lam", "shape", "rate", "phi", "phi_cov", "z"]


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    return np.allclose(a, b, rtol=rtol, atol=1e-7 * scale, equal_nan=True)


def oracle_fit(psc, stim, trace=None, **opts):
    from oracle import caviar as oc
    pr = oc.default_priors(stim.shape[0])
    return oc.caviar(psc, stim, pr["mu"], pr["beta"], pr["shape"], pr["rate"], pr["phi"], pr["phi_cov"],
                     form="reduced", trace=trace, **opts)


def gpu_fit(psc, stim, **opts):
    from circuitmap_b200 import Model
    m = Model(stim.shape[0])
    m.fit(psc, stim, method="caviar", fit_options=opts)
    return m


@pytest.fixture(scope="module")
def tiny():
    from oracle import simulate as osim
    np.random.seed(3)
    return osim.simulate(N=32, trials=300, H=4, connection_prob=0.15)


def test_tiny_map_matches_oracle_every_iteration(tiny):
    opts = dict(iters=30, seed=1, msrmp=0.4)
    tr = {"decisions": []}
    ref = oracle_fit(tiny["psc"], tiny["stim_matrix"], trace=tr, **opts)
    m = gpu_fit(tiny["psc"], tiny["stim_matrix"], save_histories=True, **opts)
    for i, nm in enumerate(NAMES):
        assert close(m.state[nm], ref[i], 1e-6), nm
    assert np.array_equal(m.state["mu"] != 0, ref[0] != 0)
    assert np.array_equal(np.nonzero(m.state["mu"])[0], np.nonzero(tiny["weights"])[0])
    for it, o in enumerate(tr["iters"]):                        # accept/reject pattern of every sweep is identical
        assert np.array_equal(m.history["lam"][it].sum(1) > 0, o["lam_sum"] > 0), it
        assert close(m.history["mu"][it], o["mu"], 1e-6) and close(m.history["rate"][it][0], o["rate"], 1e-6)
    K = tiny["psc"].shape[0]
    assert m.history["shape"].shape == (30, K) and m.history["lam"].shape == (30, 32, K)    # caviar.py:57-64
    assert m.state["receptive_fields"].shape == () and m.trial_count == K and m.time > 0
    assert m.state["lam"].dtype == np.float64 and m.state["phi_cov"].shape == (32, 2, 2)


@pytest.mark.parametrize("seed", [0, 1])
def test_c1_map_matches_oracle(seed):
    """BASELINE.json configs[0]: N=100, K=2000 trials x 900 samples, 50 iterations."""
    from oracle import simulate as osim
    np.random.seed(seed)
    sim = osim.simulate(N=100, trials=2000, H=10, connection_prob=0.1)
    opts = dict(iters=50, seed=1, msrmp=0.4)
    ref = oracle_fit(sim["psc"], sim["stim_matrix"], **opts)
    m = gpu_fit(sim["psc"], sim["stim_matrix"], **opts)
    assert np.array_equal(m.state["mu"] != 0, ref[0] != 0)
    for i, nm in enumerate(NAMES):
        assert close(m.state[nm], ref[i]), nm
    assert np.all(m.state["lam"][sim["stim_matrix"] == 0] == 0)


def test_batched_fits_equal_single_fits_bitwise():
    import torch
    from oracle import simulate as osim
    from circuitmap_b200 import optimise
    sims = [osim.simulate_fast(N=48, K=400, H=6, seed=s) for s in range(3)]
    powers = np.unique(sims[0]["stim_matrix"])[1:]
    dev = "cuda"
    f64 = dict(dtype=torch.float64, device=dev)

    def priors(B, N):
        cov = torch.zeros(B, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
        phi = torch.stack([0.1 * torch.ones(B, N, **f64), 5 * torch.ones(B, N, **f64)], -1).contiguous()
        return torch.zeros(B, N, **f64), 10 * torch.ones(B, N, **f64), 1.0, 0.1, phi, cov

    stim = torch.from_numpy(np.stack([s["stim_matrix"] for s in sims])).to(dev)
    psc = torch.from_numpy(np.stack([s["psc"] for s in sims])).to(dev)
    both = optimise.caviar_batched(stim, powers, *priors(3, 48), psc=psc, seeds=[5, 6, 7], iters=12)
    optimise.check_status(both)
    for b in range(3):
        one = optimise.caviar_batched(stim[b:b + 1].contiguous(), powers, *priors(1, 48), psc=psc[b:b + 1].contiguous(),
                                      seeds=[5 + b], iters=12)
        for nm in NAMES:
            assert torch.equal(one[nm][0], both[nm][b]), (b, nm)
    again = optimise.caviar_batched(stim, powers, *priors(3, 48), psc=psc, seeds=[5, 6, 7], iters=12)
    for nm in NAMES:
        assert torch.equal(again[nm], both[nm]), nm                 # run-to-run determinism


def test_edge_cases_against_oracle():
    """Never-stimulated neuron, all-masked trials, single power, N and K not multiples of 32, fp32 inputs."""
    from oracle import simulate as osim
    sim = osim.simulate_fast(N=37, K=333, H=5, seed=4, powers=(50,))
    stim, psc = sim["stim_matrix"].copy(), sim["psc"].copy()
    stim[5, :] = 0.0                                              # neuron 5 is never targeted
    psc[::7] *= 1e-4                                              # these trials fall under y_xcorr_thresh (lam_mask = 0)
    assert np.unique(stim).size == 2
    opts = dict(iters=10, seed=2, msrmp=0.3, minimum_spike_count=2)
    ref = oracle_fit(psc, stim, **opts)
    m = gpu_fit(psc, stim, **opts)
    for i, nm in enumerate(NAMES):
        assert close(m.state[nm], ref[i], 1e-6), nm
    assert m.state["mu"][5] == 0 and np.all(m.state["lam"][5] == 0)
    assert np.all(m.state["lam"][:, ::7] == 0)
    m32 = gpu_fit(psc.astype(np.float32), stim.astype(np.float32), **opts)
    ref32 = oracle_fit(psc.astype(np.float32), stim.astype(np.float32), **opts)
    for i, nm in enumerate(NAMES):
        assert close(m32.state[nm], ref32[i], 1e-6), nm


def test_many_powers_and_reconnection():
    """Five distinct powers (templated P>4 path); fn_scan on/off."""
    from oracle import simulate as osim
    sim = osim.simulate_fast(N=40, K=600, H=4, seed=9, powers=(30, 40, 50, 60, 70))
    for fn_scan in (True, False):
        opts = dict(iters=25, seed=3, msrmp=0.4, fn_scan=fn_scan)
        tr = {"decisions": []}
        ref = oracle_fit(sim["psc"], sim["stim_matrix"], trace=tr, **opts)
        m = gpu_fit(sim["psc"], sim["stim_matrix"], **opts)
        assert np.array_equal(m.state["mu"] != 0, ref[0] != 0)
        for i, nm in enumerate(NAMES):
            assert close(m.state[nm], ref[i], 1e-6), (fn_scan, nm)


def test_errors_are_loud():
    from circuitmap_b200 import Model
    stim = np.zeros((4, 50)); stim[0, :10] = 45.0
    psc = np.random.default_rng(0).random((50, 900))
    with pytest.raises(TypeError):
        Model(4).fit(psc, stim, fit_options={"minimax_spk_prob": 0.3})
    stim17 = np.zeros((20, 60))
    for p in range(17):
        stim17[p, p] = 10.0 + p
    with pytest.raises(RuntimeError, match="distinct stimulus powers"):
        Model(20).fit(np.ones((60, 900)), stim17)
    bad = stim.copy(); bad[1, 3] = -5.0
    with pytest.raises(RuntimeError, match="negative"):
        Model(4).fit(psc, bad)


def test_demix_to_fit_handoff_matches_two_step():
    """cm_nwd_forward's (y, ss) epilogue feeds cm_caviar_fit without re-reading the K x 900 array."""
    import os
    import torch
    from oracle import simulate as osim
    from circuitmap_b200 import NeuralDemixer, optimise
    from tests.conftest import GOLDEN
    sim = osim.simulate_fast(N=40, K=500, H=5, seed=12)
    dem = NeuralDemixer(path=os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz"))
    psc = torch.from_numpy(sim["psc"]).cuda()
    out, y, ss = dem.forward_device(psc, stats=True)
    stim = torch.from_numpy(sim["stim_matrix"]).cuda()[None].contiguous()
    f64 = dict(dtype=torch.float64, device="cuda")
    cov = torch.zeros(1, 40, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
    phi = torch.stack([0.1 * torch.ones(1, 40, **f64), 5 * torch.ones(1, 40, **f64)], -1).contiguous()
    pri = (torch.zeros(1, 40, **f64), 10 * torch.ones(1, 40, **f64), 1.0, 0.1, phi, cov)
    powers = np.unique(sim["stim_matrix"])[1:]
    a = optimise.caviar_batched(stim, powers, *pri, psc=out[None].contiguous(), seeds=[1], iters=15, msrmp=0.4)
    b = optimise.caviar_batched(stim, powers, *pri, y=y[None], ss=ss[None], seeds=[1], iters=15, msrmp=0.4)
    for nm in NAMES:
        assert torch.allclose(a[nm], b[nm], rtol=1e-6, atol=1e-10), nm   # y differs by summation order (1e-16) between the two epilogues


def test_full_size_single_fit_properties():
    """BASELINE.json configs[2] shape (N=1000, K=10000, H=10, 50 iters): size-independent properties."""
    import torch
    from oracle import simulate as osim
    from circuitmap_b200 import optimise
    N, K = 1000, 10000
    sim = osim.simulate_fast(N=N, K=K, H=10, seed=0, dtype=np.float32)
    f64 = dict(dtype=torch.float64, device="cuda")
    cov = torch.zeros(1, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
    phi = torch.stack([0.1 * torch.ones(1, N, **f64), 5 * torch.ones(1, N, **f64)], -1).contiguous()
    stim = torch.from_numpy(sim["stim_matrix"]).cuda()[None].contiguous()
    out = optimise.caviar_batched(stim, np.unique(sim["stim_matrix"])[1:], torch.zeros(1, N, **f64),
                                  10 * torch.ones(1, N, **f64), 1.0, 0.1, phi, cov,
                                  psc=torch.from_numpy(sim["psc"]).cuda()[None].contiguous(), seeds=[1], iters=50, msrmp=0.4)
    optimise.check_status(out)
    lam, mu = out["lam"][0], out["mu"][0]
    assert torch.all(lam[stim[0] == 0] == 0)                       # supp(lam) within supp(stim)
    assert out["shape"][0].item() == 1.0 + K / 2                   # caviar.py:241
    assert torch.all((lam.sum(1) > 0) == (mu != 0)) or torch.all(lam.sum(1)[mu != 0] > 0)
    assert torch.all((lam >= 0) & (lam <= 1)) and torch.isfinite(out["phi"]).all()
    truth = set(np.nonzero(sim["weights"])[0]); got = set(np.nonzero(mu.cpu().numpy())[0])
    assert len(got - truth) <= 2 and len(truth & got) >= 0.75 * len(truth)     # recovers the planted connectivity


def test_two_ctas_per_sm_variant_matches_single_fits():
    """Batches with >= 2 fits per SM run the 8-warp CTA variant of the persistent kernel (two CTAs per SM).
    Its results must agree with the 16-warp variant used for small batches (different reduction widths -> not bitwise)."""
    import torch
    from oracle import simulate as osim
    from circuitmap_b200 import optimise
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    B = 2 * sms
    sims = [osim.simulate_fast(N=40, K=320, H=5, seed=s) for s in range(4)]
    powers = np.unique(sims[0]["stim_matrix"])[1:]
    f64 = dict(dtype=torch.float64, device="cuda")
    stim4 = torch.from_numpy(np.stack([s["stim_matrix"] for s in sims])).cuda()
    psc4 = torch.from_numpy(np.stack([s["psc"] for s in sims])).cuda()
    idx = torch.arange(B, device="cuda") % 4

    def priors(b):
        cov = torch.zeros(b, 40, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
        phi = torch.stack([0.1 * torch.ones(b, 40, **f64), 5 * torch.ones(b, 40, **f64)], -1).contiguous()
        return torch.zeros(b, 40, **f64), 10 * torch.ones(b, 40, **f64), 1.0, 0.1, phi, cov

    big = optimise.caviar_batched(stim4[idx].contiguous(), powers, *priors(B), psc=psc4[idx].contiguous(),
                                  seeds=[7] * B, iters=14, msrmp=0.4)
    optimise.check_status(big)
    small = optimise.caviar_batched(stim4, powers, *priors(4), psc=psc4, seeds=[7] * 4, iters=14, msrmp=0.4)
    for b in (0, 1, 2, 3, 4, B - 1):
        for nm in NAMES:
            assert torch.equal(big["mu"][b] != 0, small["mu"][b % 4] != 0)
            assert torch.allclose(big[nm][b], small[nm][b % 4], rtol=1e-6, atol=1e-9), (b, nm)


def test_c4_batched_sweep_shape():
    """BASELINE.json configs[3] shape: a batch of independent maps with N=500, K=5000 (here 2 fits per SM of distinct
    seeds over 6 distinct maps, so the two-CTAs-per-SM variant runs).  One map is checked against the oracle run for
    the full 50 iterations; all of them against the size-independent properties and against their own single fit."""
    import torch
    from oracle import caviar as oc, simulate as osim
    from circuitmap_b200 import optimise
    N, K, nmaps = 500, 5000, 6
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    B = 2 * sms
    sims = [osim.simulate_fast(N=N, K=K, H=10, seed=40 + s) for s in range(nmaps)]
    powers = np.unique(sims[0]["stim_matrix"])[1:]
    f64 = dict(dtype=torch.float64, device="cuda")
    stim_m = torch.from_numpy(np.stack([s["stim_matrix"] for s in sims])).cuda()
    psc_m = torch.from_numpy(np.stack([s["psc"] for s in sims])).float().cuda()
    idx = torch.arange(B, device="cuda") % nmaps

    def priors(b):
        cov = torch.zeros(b, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
        phi = torch.stack([0.1 * torch.ones(b, N, **f64), 5 * torch.ones(b, N, **f64)], -1).contiguous()
        return torch.zeros(b, N, **f64), 10 * torch.ones(b, N, **f64), 1.0, 0.1, phi, cov

    seeds = [1 + b for b in range(B)]
    out = optimise.caviar_batched(stim_m[idx].contiguous(), powers, *priors(B), psc=psc_m[idx].contiguous(), seeds=seeds,
                                  iters=50, msrmp=0.4, want_lam=False)
    optimise.check_status(out)
    assert torch.all(out["shape"] == 1.0 + K / 2) and torch.isfinite(out["phi"]).all() and torch.isfinite(out["mu"]).all()
    for b in (0, 1, B - 1):
        truth = set(np.nonzero(sims[b % nmaps]["weights"])[0]); got = set(np.nonzero(out["mu"][b].cpu().numpy())[0])
        assert len(got - truth) <= 2 and len(truth & got) >= 0.7 * len(truth)
    # fit 7 alone (16-warp CTA variant) and in the oracle
    b = 7
    m = b % nmaps
    one = optimise.caviar_batched(stim_m[m:m + 1].contiguous(), powers, *priors(1), psc=psc_m[m:m + 1].contiguous(),
                                  seeds=[seeds[b]], iters=50, msrmp=0.4)
    assert torch.equal(one["mu"][0] != 0, out["mu"][b] != 0)
    for nm in ("mu", "beta", "shape", "rate", "phi", "phi_cov", "z"):
        assert torch.allclose(one[nm][0], out[nm][b], rtol=1e-6, atol=1e-9), nm
    pr = oc.default_priors(N)
    ref = oc.caviar(psc_m[m].double().cpu().numpy(), sims[m]["stim_matrix"], pr["mu"], pr["beta"], pr["shape"], pr["rate"],
                    pr["phi"], pr["phi_cov"], iters=50, seed=seeds[b], msrmp=0.4)
    assert np.array_equal(one["mu"][0].cpu().numpy() != 0, ref[0] != 0)          # identical connected set
    for i, nm in enumerate(["mu", "beta", "lam", "shape", "rate", "phi", "phi_cov", "z"]):
        a_, b_ = one[nm][0].cpu().numpy(), np.asarray(ref[i], float)
        assert np.allclose(a_, b_, rtol=1e-4, atol=1e-7 * max(np.max(np.abs(b_)), 1e-300), equal_nan=True), nm


def test_c5_large_single_map_shape():
    """BASELINE.json configs[4] shape (N=5000, K=100000, H=10, 50 iterations) as ONE fit on one GPU: size-independent
    properties (the oracle needs hours at this size) and recovery of the planted connectivity."""
    import torch
    from oracle import simulate as osim
    from circuitmap_b200 import optimise
    N, K = 5000, 100000
    sim = osim.simulate_fast(N=N, K=K, H=10, seed=0, dtype=np.float32)
    f64 = dict(dtype=torch.float64, device="cuda")
    cov = torch.zeros(1, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
    phi = torch.stack([0.1 * torch.ones(1, N, **f64), 5 * torch.ones(1, N, **f64)], -1).contiguous()
    stim = torch.from_numpy(sim["stim_matrix"]).cuda()[None].contiguous()
    out = optimise.caviar_batched(stim, np.unique(sim["stim_matrix"])[1:], torch.zeros(1, N, **f64),
                                  10 * torch.ones(1, N, **f64), 1.0, 0.1, phi, cov,
                                  psc=torch.from_numpy(sim["psc"]).cuda()[None].contiguous(), seeds=[1], iters=50, msrmp=0.4)
    optimise.check_status(out)
    lam, mu = out["lam"][0], out["mu"][0]
    assert not torch.any(lam[stim[0] == 0] != 0)                   # supp(lam) within supp(stim)
    assert out["shape"][0].item() == 1.0 + K / 2                   # caviar.py:241
    assert torch.all(lam.sum(1)[mu != 0] > 0)
    assert torch.all((lam >= 0) & (lam <= 1)) and torch.isfinite(out["phi"]).all() and torch.isfinite(out["rate"]).all()
    truth = set(np.nonzero(sim["weights"])[0]); got = set(np.nonzero(mu.cpu().numpy())[0])
    assert len(got - truth) <= 5 and len(truth & got) >= 0.7 * len(truth)


def test_chain_teams_pass_jobs_and_inverse_variants():
    """Round-2 rearrangements of a single large fit (csrc/caviar_fit.inl): the second chain team (rows without a common
    trial overlap, sweep_chain_fast) and the O(K) / O(nnz) pass jobs on the helper CTAs must not change a bit -- every
    entry of the running prediction sees the same updates in the same order, every map output is computed by one thread
    as before.  The recursive tile inverse sums in a different order than the column-by-column substitution it
    replaces: same connected set, values to rounding.  Diagnostics bits 10 / 11 / 12 of cm_caviar_debug_phase_cycles
    switch the old forms back on."""
    import torch
    from oracle import simulate as osim
    from circuitmap_b200 import optimise, _lib
    lib = _lib.load()
    N, K = 700, 6000
    sim = osim.simulate_fast(N=N, K=K, H=10, seed=77)
    f64 = dict(dtype=torch.float64, device="cuda")
    cov = torch.zeros(1, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
    phi = torch.stack([0.1 * torch.ones(1, N, **f64), 5 * torch.ones(1, N, **f64)], -1).contiguous()
    stim = torch.from_numpy(sim["stim_matrix"][None]).cuda()
    psc = torch.from_numpy(sim["psc"][None]).cuda()
    args = (stim, np.unique(sim["stim_matrix"])[1:], torch.zeros(1, N, **f64), 10 * torch.ones(1, N, **f64), 1.0, 0.1, phi, cov)
    outs = {}
    try:
        for name, bits in (("default", 0), ("one_team", 1 << 11), ("no_pass_jobs", 1 << 10), ("column_inverse", 1 << 12)):
            assert lib.cm_caviar_debug_phase_cycles(None, 0, bits) == 0
            outs[name] = optimise.caviar_batched(*args, psc=psc, seeds=[3], iters=14, msrmp=0.4)
            optimise.check_status(outs[name])
    finally:
        lib.cm_caviar_debug_phase_cycles(None, 0, 0)
    ref = outs["default"]
    assert int((ref["mu"][0] != 0).sum()) > 20
    for nm in NAMES:
        assert torch.equal(ref[nm], outs["one_team"][nm]), nm
        assert torch.equal(ref[nm], outs["no_pass_jobs"][nm]), nm
    alt = outs["column_inverse"]
    assert torch.equal(ref["mu"] != 0, alt["mu"] != 0)
    for nm in NAMES:
        a, b = ref[nm].double(), alt[nm].double()
        assert torch.allclose(a, b, rtol=1e-7, atol=1e-9 * float(b.abs().max()), equal_nan=True), nm


def test_panel_gemm_helper_ctas_are_bitwise_neutral():
    """A single large fit gets helper CTAs for the column tiles of its panel GEMMs (csrc/caviar_fit.inl,
    panel_gemm_dist).  Every output element is still computed by one warp in the same order: the fit must be bitwise
    identical with and without helpers (CM_CAVIAR_HELPERS=0), also for two fits side by side."""
    import torch
    from oracle import simulate as osim
    from circuitmap_b200 import optimise
    N, K, B = 700, 6000, 2
    sims = [osim.simulate_fast(N=N, K=K, H=10, seed=70 + s) for s in range(B)]
    f64 = dict(dtype=torch.float64, device="cuda")
    cov = torch.zeros(B, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
    phi = torch.stack([0.1 * torch.ones(B, N, **f64), 5 * torch.ones(B, N, **f64)], -1).contiguous()
    stim = torch.from_numpy(np.stack([s["stim_matrix"] for s in sims])).cuda()
    psc = torch.from_numpy(np.stack([s["psc"] for s in sims])).cuda()
    args = (stim, np.unique(sims[0]["stim_matrix"])[1:], torch.zeros(B, N, **f64), 10 * torch.ones(B, N, **f64), 1.0, 0.1,
            phi, cov)
    old = os.environ.get("CM_CAVIAR_HELPERS")
    try:
        os.environ["CM_CAVIAR_HELPERS"] = "0"
        ref = optimise.caviar_batched(*args, psc=psc, seeds=[1, 2], iters=12, msrmp=0.4)
        os.environ["CM_CAVIAR_HELPERS"] = "15"
        out = optimise.caviar_batched(*args, psc=psc, seeds=[1, 2], iters=12, msrmp=0.4)
        os.environ["CM_CAVIAR_HELPERS"] = "5"
        out5 = optimise.caviar_batched(*args, psc=psc, seeds=[1, 2], iters=12, msrmp=0.4)
    finally:
        if old is None:
            os.environ.pop("CM_CAVIAR_HELPERS", None)
        else:
            os.environ["CM_CAVIAR_HELPERS"] = old
    for o in (ref, out, out5):
        optimise.check_status(o)
    for nm in NAMES:
        assert torch.equal(ref[nm], out[nm]), nm
        assert torch.equal(ref[nm], out5[nm]), nm
