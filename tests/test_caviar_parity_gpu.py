"""GPU parity, round 2: the configurations the benchmark quotes, the C1 matrix of BASELINE.md section 4, decision lists with
their margins, demixer-precision -> decision stability, multi-event reconnection, reference-executed fixtures.

Everything goes through the C ABI (Model.fit / optimise.caviar / caviar_batched).  Tolerances as BASELINE.json states them:
connected set and every accept/reject decision IDENTICAL, posteriors within 1e-4 relative (fp64 kernel: measured ~1e-8).
"""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu

RTOL = 1e-4
NAMES = ["mu", "beta", "lam", "shape", "rate", "phi", "phi_cov", "z"]


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(float(np.nanmax(np.abs(b))) if b.size and np.isfinite(b).any() else 0.0, 1e-300)
    return np.allclose(a, b, rtol=rtol, atol=1e-7 * scale, equal_nan=True)


def oracle_fit(psc, stim, trace=None, **opts):
    from oracle import caviar as oc
    pr = oc.default_priors(stim.shape[0])
    with np.errstate(all="ignore"):
        return oc.caviar(psc, stim, pr["mu"], pr["beta"], pr["shape"], pr["rate"], pr["phi"], pr["phi_cov"],
                         form="reduced", trace=trace, **opts)


def gpu_fit(psc, stim, **opts):
    from circuitmap_b200 import Model
    m = Model(stim.shape[0])
    m.fit(psc, stim, method="caviar", fit_options=opts)
    return m


def decision_margin(trace, minspk=3.0):
    """Smallest distance of any gated decision of the oracle run from flipping (SURVEY 7 hard part 4)."""
    gated = [d for d in trace["decisions"] if d[0] > 1]
    if not gated:
        return np.inf
    return min(min(abs(d[3] - d[5]) for d in gated), min(abs(d[4] - minspk) for d in gated))


def assert_same_decisions(m_hist_lam, m_hist_mu, trace, iters, N):
    """The GPU's accept/reject of every (iteration, neuron) equals the oracle's decision list."""
    ok = np.ones((iters, N), dtype=bool)
    for (it, _, n, _, _, _, okv) in trace["decisions"]:
        ok[it, n] = okv == 1.0
    got = (np.asarray(m_hist_lam) != 0).any(2)
    # a row that was accepted can still be all-zero only if every estimate underflowed -- never at these sizes;
    # a rejected row is always zero (caviar.py:224)
    assert not np.any(got & ~ok), "GPU kept a row the oracle rejected"
    want = np.stack([t["lam_sum"] > 0 for t in trace["iters"]])
    assert np.array_equal(got, want)


# ------------------------------------------------------------------------------------------------ C1 matrix
_C1 = {}


def c1_sim(seed):
    if seed not in _C1:
        from oracle import simulate as osim
        from circuitmap_b200 import NeuralDemixer
        np.random.seed(seed)
        sim = osim.simulate(N=100, trials=2000, H=10, connection_prob=0.1)
        dem = NeuralDemixer(path=os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz"))
        sim["psc_dem"] = dem(sim["psc"], verbose=False)                       # fp32 network = the reference's arithmetic
        _C1[seed] = sim
    return _C1[seed]


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("which", ["raw", "demixed"])
@pytest.mark.parametrize("msrmp", [0.3, 0.4])
def test_c1_matrix(seed, which, msrmp):
    """BASELINE.json configs[0] / BASELINE.md section 4: seeds 0-2 x {raw, NWD-demixed} x {default msrmp 0.3, 0.4 as
    scripts/run_simulations.py:54-61}; 50 iterations.  Connected set, every decision, all eight state keys."""
    sim = c1_sim(seed)
    psc = sim["psc"] if which == "raw" else sim["psc_dem"]
    opts = dict(iters=50, seed=1) if msrmp == 0.3 else dict(iters=50, seed=1, msrmp=0.4)
    tr = {"decisions": []}
    ref = oracle_fit(psc, sim["stim_matrix"], trace=tr, **opts)
    m = gpu_fit(psc, sim["stim_matrix"], save_histories=True, **opts)
    margin = decision_margin(tr)
    assert margin > 1e-7, margin                   # fp64 on both sides: anything above ~1e-10 is decided identically
    assert np.array_equal(m.state["mu"] != 0, ref[0] != 0)
    assert_same_decisions(m.history["lam"], m.history["mu"], tr, 50, 100)
    for i, nm in enumerate(NAMES):
        assert close(m.state[nm], ref[i]), (nm, margin)
    print("C1 seed %d %s msrmp %.1f: min decision margin %.3g, connected %d" % (seed, which, msrmp, margin,
                                                                              int((ref[0] != 0).sum())))


def test_c1_reference_fixture_direct():
    """GPU against the fixture the UNMODIFIED reference produced through the jax shim (tests/golden/caviar_ref_C1_seed0.npz)."""
    g = np.load(os.path.join(GOLDEN, "caviar_ref_C1_seed0.npz"))
    sim = c1_sim(0)
    assert np.array_equal(sim["stim_matrix"], g["stim"].astype(np.float64))
    m = gpu_fit(sim["psc"], sim["stim_matrix"], iters=50, seed=1, msrmp=0.4, save_histories=True)
    assert np.array_equal(m.state["mu"] != 0, g["mu"] != 0)
    for nm in ["mu", "beta", "shape", "rate", "phi", "phi_cov", "z"]:
        assert close(m.state[nm], g[nm]), nm
    assert close(m.state["lam"][sim["stim_matrix"] > 0], g["lam_on_support"])
    assert np.array_equal((m.history["lam"] != 0).any(2), g["hist_lam_rowany"])       # all 5000 decisions
    assert close(m.history["mu"], g["hist_mu"]) and close(m.history["rate"][:, 0], g["hist_rate"])


def test_tiny_reference_fixture_direct():
    g = np.load(os.path.join(GOLDEN, "caviar_ref_tiny_N32_K300.npz"))
    m = gpu_fit(g["psc"], g["stim"].astype(np.float64), iters=30, seed=1, msrmp=0.4, save_histories=True)
    for nm in NAMES:
        assert close(m.state[nm], g[nm], 1e-5), nm
        assert close(m.history[nm] if nm not in ("shape", "rate") else m.history[nm], g["hist_" + nm], 1e-5), nm
    assert np.array_equal((m.history["lam"] != 0).any(2), (g["hist_lam"] != 0).any(2))


# ------------------------------------------------------------------------------------------------ C3: the benchmarked config
def test_c3_full_size_fit_matches_oracle():
    """BASELINE.json configs[2] -- the shape bench.py quotes (N=1000, K=10000, H=10, 50 iterations): identical connected
    set, 1e-4 on all eight state keys, identical decisions in the gated iterations."""
    from oracle import simulate as osim
    sim = osim.simulate_fast(N=1000, K=10000, H=10, seed=0, dtype=np.float32)
    psc, stim = sim["psc"], sim["stim_matrix"].astype(np.float64)
    opts = dict(iters=50, seed=1, msrmp=0.4)
    tr = {"decisions": []}
    ref = oracle_fit(psc.astype(np.float64), stim, trace=tr, **opts)
    m = gpu_fit(psc, stim, **opts)
    margin = decision_margin(tr)
    assert margin > 1e-8, margin
    assert np.array_equal(m.state["mu"] != 0, ref[0] != 0)
    for i, nm in enumerate(NAMES):
        assert close(m.state[nm], ref[i]), (nm, margin)
    truth = set(np.nonzero(sim["weights"])[0]); got = set(np.nonzero(m.state["mu"])[0])
    assert len(got - truth) <= 2 and len(truth & got) >= 0.75 * len(truth)
    print("C3: %d connected (%d planted), min decision margin %.3g over %d decisions" % (
        len(got), len(truth), margin, len(tr["decisions"])))


# ------------------------------------------------------------------------------------------------ demixer precision -> decisions
@pytest.mark.parametrize("seed", [0, 1])
def test_fp16_demixer_does_not_move_caviar_decisions(seed):
    """the fp16 tensor-core demixer (the throughput bench.py quotes) feeds CAVIaR the same lam_mask
    (sum x^2 > 1e-2, caviar.py:30) and the same connected set as the fp32 demixer on C1."""
    from circuitmap_b200 import NeuralDemixer
    sim = c1_sim(seed)
    dem16 = NeuralDemixer(path=os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz"), precision="fp16")
    p16 = dem16(sim["psc"], verbose=False)
    p32 = sim["psc_dem"]
    assert np.array_equal(np.sum(p16 * p16, 1) > 1e-2, np.sum(p32 * p32, 1) > 1e-2)        # lam_mask unchanged
    for msrmp in (0.3, 0.4):
        a = gpu_fit(p32, sim["stim_matrix"], iters=50, seed=1, msrmp=msrmp)
        b = gpu_fit(p16, sim["stim_matrix"], iters=50, seed=1, msrmp=msrmp)
        assert np.array_equal(a.state["mu"] != 0, b.state["mu"] != 0), msrmp               # same connected set
        con = a.state["mu"] != 0
        assert np.allclose(a.state["mu"][con], b.state["mu"][con], rtol=2e-2), msrmp        # weights move with the traces


# ------------------------------------------------------------------------------------------------ reconnection
def reconnect_map(seed, N=24, H=3, reps=20, spont_p=0.2):
    """Constructed experiment for reconnect_spont_cells (caviar.py:102-144): a high spontaneous rate raises the in-loop
    threshold msrmp + spont_rate, weakly-responding cells are pruned and their events land in z; two 'rare' cells are
    stimulated on 2 / 3 trials only (single-event candidates).  Traces are synthetic spikes with trapz = y."""
    rng = np.random.default_rng(seed)
    powers = np.array([45., 55., 65.])
    cols = []
    for r in range(reps):
        for p in powers:
            order = rng.permutation(N - 2)
            for h in range((N - 2) // H):
                col = np.zeros(N); col[order[h * H:(h + 1) * H]] = p; cols.append(col)
    stim = np.array(cols).T
    extra = []
    for n, pw in [(N - 2, [65., 65.]), (N - 1, [65., 55., 45.])]:
        for p in pw:
            col = np.zeros(N); col[n] = p; extra.append(col)
    stim = np.concatenate([stim, np.array(extra).T], 1)
    stim = stim[:, rng.permutation(stim.shape[1])]
    K = stim.shape[1]
    w = np.zeros(N); prob = np.zeros(N)
    w[0:3] = rng.uniform(20, 35, 3); prob[0:3] = 0.9
    w[3:9] = rng.uniform(10, 25, 6); prob[3:9] = rng.uniform(0.33, 0.42, 6)
    w[N - 2] = 15.; prob[N - 2] = 0.5
    w[N - 1] = 12.; prob[N - 1] = 0.5
    spk = (rng.random((N, K)) < prob[:, None]) & (stim > 0)
    y = (w[:, None] * spk * rng.lognormal(0, 0.05, (N, K))).sum(0)
    y = y + (rng.random(K) < spont_p) * rng.uniform(7, 25, K) + rng.normal(0, 0.05, K)
    psc = np.zeros((K, 900)); psc[:, 100] = y; psc[:, 200] = 1.0; psc[:, 201] = -1.0
    return psc, np.ascontiguousarray(stim)


@pytest.mark.parametrize("seed,msc,want_events,want_nan", [(16, 1, 3, 2), (6, 1, 2, 2), (1, 1, 2, 0), (16, 3, 1, 0),
                                                           (9, 1, 2, 1)])
def test_reconnect_sequences_match_oracle(seed, msc, want_events, want_nan):
    """Several cells reconnect in sequence (events of earlier ones are removed from z), first-arg-max ties between
    single-event cells, sem() of a single sample = NaN."""
    psc, stim = reconnect_map(seed)
    opts = dict(iters=30, seed=1, msrmp=0.3, minimum_spike_count=msc)
    tr = {"decisions": []}
    ref = oracle_fit(psc, stim, trace=tr, **opts)
    assert len(tr["reconnect"]) == want_events and int(np.isnan(ref[1]).sum()) == want_nan     # the case is what it claims
    m = gpu_fit(psc, stim, **opts)
    assert np.array_equal(m.state["mu"] != 0, ref[0] != 0)
    assert np.array_equal(np.isnan(m.state["beta"]), np.isnan(ref[1]))
    for i, nm in enumerate(NAMES):
        assert close(m.state[nm], ref[i], 1e-6), nm
    for f, _, _ in tr["reconnect"]:
        assert m.state["mu"][f] != 0 and np.all(np.isin(m.state["lam"][f], [0.0, 1.0]))      # lam[focus, locs] = 1


# ------------------------------------------------------------------------------------------------ boundary variants
def test_uint8_codes_csr_output_and_device_inputs_are_bitwise_the_float_path():
    import torch
    from oracle import simulate as osim
    from circuitmap_b200 import Model, NeuralDemixer, optimise
    sim = osim.simulate_fast(N=48, K=400, H=6, seed=21)
    stim, psc = sim["stim_matrix"], sim["psc"]
    f64 = dict(dtype=torch.float64, device="cuda")
    cov = torch.zeros(1, 48, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
    phi = torch.stack([0.1 * torch.ones(1, 48, **f64), 5 * torch.ones(1, 48, **f64)], -1).contiguous()
    pri = (torch.zeros(1, 48, **f64), 10 * torch.ones(1, 48, **f64), 1.0, 0.1, phi, cov)
    sd, pd = torch.from_numpy(stim).cuda()[None].contiguous(), torch.from_numpy(psc).cuda()[None].contiguous()
    nnz, values = optimise.scan_stim(sd)
    assert nnz == np.count_nonzero(stim) and np.array_equal(values, np.unique(stim)[1:])
    powers = optimise.powers_like_reference(nnz, values, stim.size)
    a = optimise.caviar_batched(sd, powers, *pri, psc=pd, seeds=[3], iters=12, msrmp=0.4)
    codes, p2, n2 = optimise.pack_stim_host(np.ascontiguousarray(stim))
    assert np.array_equal(p2, powers) and n2 == nnz
    b = optimise.caviar_batched(codes.cuda()[None].contiguous(), powers, *pri, psc=pd, seeds=[3], iters=12, msrmp=0.4,
                                want_lam=False, lam_csr=True, nnz_cap=nnz)
    optimise.check_status(a); optimise.check_status(b)
    for nm in ("mu", "beta", "shape", "rate", "phi", "phi_cov", "z"):
        assert torch.equal(a[nm], b[nm]), nm
    csr = optimise.CsrLam(b["lam_csr_val"][0], b["lam_csr_col"][0], b["lam_csr_ptr"][0], 400)
    assert np.array_equal(csr.toarray(), a["lam"][0].cpu().numpy())
    # the drop-in classes with device tensors: NeuralDemixer(tensor) -> tensor carrying (y, ss) -> Model.fit(tensor, tensor)
    dem = NeuralDemixer(path=os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz"))
    d_dev = dem(torch.from_numpy(psc).cuda(), verbose=False)
    assert d_dev.is_cuda and d_dev.cm_y.shape == (400,)
    d_host = dem(psc, verbose=False)
    assert np.array_equal(d_dev.cpu().numpy(), d_host)
    m_dev, m_host = Model(48), Model(48)
    m_dev.fit(d_dev, torch.from_numpy(stim).cuda(), fit_options=dict(iters=12, seed=3, msrmp=0.4))
    m_host.fit(d_host, stim, fit_options=dict(iters=12, seed=3, msrmp=0.4))
    assert np.array_equal(m_dev.state["mu"] != 0, m_host.state["mu"] != 0)
    for nm in NAMES:
        assert close(m_dev.state[nm], m_host.state[nm], 1e-9), nm       # y from the demixer epilogue vs from the prologue: 1e-16


def test_batch_larger_than_one_resident_wave_uses_the_fit_queue():
    """B = 2 * SMs + 7 fits: more than one resident wave of the 8-warp variant -> CTAs pull fits from the device queue.
    Every fit equals its single-fit run (same variant: bitwise)."""
    import torch
    from oracle import simulate as osim
    from circuitmap_b200 import optimise
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    B = 2 * sms + 7
    sims = [osim.simulate_fast(N=40, K=320, H=5, seed=s) for s in range(5)]
    powers = np.unique(sims[0]["stim_matrix"])[1:]
    f64 = dict(dtype=torch.float64, device="cuda")
    stim5 = torch.from_numpy(np.stack([s["stim_matrix"] for s in sims])).cuda()
    psc5 = torch.from_numpy(np.stack([s["psc"] for s in sims])).cuda()
    idx = torch.arange(B, device="cuda") % 5

    def priors(b):
        cov = torch.zeros(b, 40, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
        phi = torch.stack([0.1 * torch.ones(b, 40, **f64), 5 * torch.ones(b, 40, **f64)], -1).contiguous()
        return torch.zeros(b, 40, **f64), 10 * torch.ones(b, 40, **f64), 1.0, 0.1, phi, cov

    seeds = [11 + (b % 13) for b in range(B)]
    big = optimise.caviar_batched(stim5[idx].contiguous(), powers, *priors(B), psc=psc5[idx].contiguous(), seeds=seeds,
                                  iters=14, msrmp=0.4)
    optimise.check_status(big)
    old = os.environ.get("CM_CAVIAR_CTA")
    os.environ["CM_CAVIAR_CTA"] = "256"
    try:
        for b in (0, 1, sms, 2 * sms - 1, 2 * sms, B - 1):
            one = optimise.caviar_batched(stim5[b % 5:b % 5 + 1].contiguous(), powers, *priors(1),
                                          psc=psc5[b % 5:b % 5 + 1].contiguous(), seeds=[seeds[b]], iters=14, msrmp=0.4)
            for nm in NAMES:
                assert torch.equal(one[nm][0], big[nm][b]), (b, nm)
    finally:
        if old is None:
            os.environ.pop("CM_CAVIAR_CTA", None)
        else:
            os.environ["CM_CAVIAR_CTA"] = old


def test_second_device_in_the_same_process():
    """timing events are per device -- a fit / demix on cuda:1 after cuda:0 in one thread must succeed."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import simulate as osim
    from circuitmap_b200 import Model, NeuralDemixer, _lib, optimise
    sim = osim.simulate_fast(N=32, K=256, H=4, seed=1)
    res = []
    for d in ("cuda:0", "cuda:1", "cuda:0"):
        dem = NeuralDemixer(path=os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz"), device=d)
        x = dem(sim["psc"], verbose=False)
        out = optimise.caviar(x, sim["stim_matrix"], np.zeros(32), 10 * np.ones(32), 1.0, 0.1,
                              np.c_[0.1 * np.ones(32), 5 * np.ones(32)],
                              np.array([[[0.1, 0], [0, 1.0]]] * 32), device=d, iters=8, seed=2)
        with torch.cuda.device(d):
            assert _lib.load().cm_last_main_kernel_ms() > 0
        res.append((x, out[0]))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][1], res[2][1])


def test_large_powers_take_the_reference_expression_of_the_monte_carlo_term():
    """mean_s log(f / (1 - f)) (caviar.py:213-215) saturates for large powers; kernel and reduced oracle then evaluate
    the reference's own expression instead of the linear shortcut; the literal oracle (the reference's structure) agrees."""
    from oracle import caviar as oc, simulate as osim
    sim = osim.simulate_fast(N=24, K=240, H=4, seed=5, powers=(60, 90, 99))
    stim = np.ascontiguousarray(sim["stim_matrix"])
    stim[stim == 90] = 160.0
    stim[stim == 99] = 320.0
    opts = dict(iters=8, seed=1, msrmp=0.4)
    pr = oc.default_priors(24)
    with np.errstate(all="ignore"):
        lit = oc.caviar(sim["psc"], stim, pr["mu"], pr["beta"], pr["shape"], pr["rate"], pr["phi"], pr["phi_cov"], form="literal", **opts)
    ref = oracle_fit(sim["psc"], stim, **opts)
    m = gpu_fit(sim["psc"], stim, **opts)
    assert np.array_equal(m.state["mu"] != 0, ref[0] != 0) and np.array_equal(m.state["mu"] != 0, lit[0] != 0)
    for i, nm in enumerate(NAMES):
        assert close(m.state[nm], ref[i], 1e-6), nm
        assert close(m.state[nm], lit[i], 1e-5), nm
