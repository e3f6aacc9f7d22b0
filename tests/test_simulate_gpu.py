"""GPU: the device generator cm_simulate (csrc/simulate.cu) against the DISTRIBUTIONS of circuitmap/simulation.py:25-215.

The reference draws from NumPy's unseeded global stream, so parity is distribution-level: design invariants hold exactly,
moments within sampling error, and a fit on a device-generated map recovers the planted connectivity exactly like a fit on
a map from the oracle's order-faithful restatement of `simulate` does."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sim100():
    from circuitmap_b200 import simulate
    return simulate(N=100, trials=2000, H=10, connection_prob=0.1, seed=11)


def test_design_invariants(sim100):
    stim, N, K, H = sim100["stim_matrix"], 100, 2000, 10
    assert stim.shape == (N, K) and stim.dtype == np.float64 and sim100["psc"].shape == (K, 900)
    assert set(np.unique(stim)) == {0.0, 45.0, 55.0, 65.0}
    assert np.all((stim > 0).sum(0) == H)                               # exactly H targets per trial (N % H == 0)
    assert np.all(np.array([np.unique(c[c > 0]).size for c in stim.T]) == 1)      # one power per trial
    per = (stim > 0).sum(1)
    assert per.min() >= K * H // N - 1 and per.max() <= K * H // N + 1             # blockwise: every cell equally often
    # every pass shows each hologram at each power: per-neuron power counts differ by at most one pass
    for p in (45.0, 55.0, 65.0):
        c = (stim == p).sum(1)
        assert c.max() - c.min() <= 1, p
    # the higher powers are filled first (simulation.py:48): the incomplete last pass favours 65 over 45
    assert (stim == 65.0).sum() >= (stim == 55.0).sum() >= (stim == 45.0).sum()
    assert np.array_equal(sim100["I"], stim.max(0))
    # trials are shuffled: the power sequence is not sorted in blocks
    assert 0.5 < np.mean(np.diff(sim100["I"]) != 0) < 0.8


def test_weights_distribution():
    import torch
    from circuitmap_b200.simulation import simulate_batch
    res = simulate_batch(list(range(64)), N=200, trials=300, H=10, connection_prob=0.1)
    w = res["weights"].cpu().numpy()
    assert np.all((w != 0).sum(1) == 20)                                 # int(connection_prob * N), exactly
    strong = (w >= 20).sum(1)
    conn = w[w != 0]
    assert conn.min() >= 9.0                                             # weak = Exp(4) + 9, strong = U[20, 40]
    weak = conn[conn < 20]
    # Exp(4) + 9 truncated below 20 (the few weak draws above 20 are counted as strong here)
    assert abs(weak.mean() - (9 + 4 - 11 * np.exp(-11 / 4) / (1 - np.exp(-11 / 4)))) < 0.25
    assert 4 * 64 <= strong.sum() <= 4 * 64 + 0.12 * 16 * 64             # ceil(0.2 * 20) = 4 strong per map (+ weak tail)
    big = conn[conn >= 20]
    assert big.max() <= 40 + 40 and abs(np.median(big) - 29.5) < 2.5
    # different seeds give different maps, the same seed the same map
    again = simulate_batch([5, 5, 6], N=200, trials=300, H=10, connection_prob=0.1)
    assert torch.equal(again["psc"][0], again["psc"][1]) and not torch.equal(again["psc"][0], again["psc"][2])
    assert torch.equal(again["codes"][0], again["codes"][1]) and torch.equal(again["weights"][0], res["weights"][5])


def test_noise_moments_and_gp_covariance():
    """Trials whose targets are all unconnected and that drew no spontaneous event hold only noise: the GP of
    simulation.py:211-215 (scale 4e-3, squared-exponential length 50) plus iid sigma 6e-4."""
    from circuitmap_b200 import simulate
    sim = simulate(N=100, trials=4000, H=10, connection_prob=0.05, seed=3, spont_prob=0.0)
    stim, psc, w = sim["stim_matrix"], sim["psc"], sim["weights"]
    quiet = np.nonzero((stim[w != 0] > 0).sum(0) == 0)[0]
    assert quiet.size > 1500
    x = psc[quiet]
    assert abs(x.mean()) < 2e-4
    var = x.var()
    assert abs(var - (4e-3 ** 2 + 6e-4 ** 2)) < 0.06 * 4e-3 ** 2
    xc = x - x.mean()
    for d in (1, 10, 25, 50, 100, 200):
        cov = np.mean(xc[:, :-d] * xc[:, d:])
        want = 4e-3 ** 2 * np.exp(-d ** 2 / (2 * 50.0 ** 2))
        assert abs(cov - want) < 0.08 * 4e-3 ** 2, (d, cov, want)
    # edges are as noisy as the middle (the filtered white noise is extended beyond both ends)
    assert abs(x[:, :20].var() / x[:, 440:460].var() - 1) < 0.25 and abs(x[:, -20:].var() / x[:, 440:460].var() - 1) < 0.25


def test_evoked_and_spontaneous_events(sim100):
    stim, psc, w = sim100["stim_matrix"], sim100["psc"], sim100["weights"]
    y = np.trapezoid(psc, axis=-1)
    conn = np.nonzero(w)[0]
    assert conn.size == 10
    # connected cells respond in >= ~40 % of their top-power trials with an area close to their weight (:98-108,17-21)
    for n in conn:
        top = np.nonzero(stim[n] == 65.0)[0]
        alone = top[(stim[conn][:, top] > 0).sum(0) == 1]              # no other connected cell in the hologram
        resp = y[alone] > 0.5 * w[n]
        assert resp.mean() >= 0.3, (n, resp.mean())
        areas = y[alone][resp]
        assert abs(np.median(areas) / w[n] - 1.0) < 0.25, (n, np.median(areas), w[n])
    # latencies: 160 + Gamma(1e4 / I^2, 15) -> responses start after sample 160
    strongest = conn[np.argmax(w[conn])]
    tr = np.nonzero((stim[strongest] == 65.0) & (y > 0.5 * w[strongest]))[0]
    onset = np.array([np.argmax(psc[k] > 0.02 * psc[k].max() + 0.02) for k in tr])
    assert onset.min() >= 150 and 180 < np.median(onset) < 260           # mean latency at 65: 160 + 2.37 * 15 = 195
    # spontaneous events: ~5 % of the trials without an evoked response still carry a PSC-sized area
    quiet = np.nonzero((stim[conn] > 0).sum(0) == 0)[0]
    frac = np.mean(y[quiet] > 5.0)
    assert 0.02 < frac < 0.09, frac


def test_fit_on_device_generated_map_recovers_connectivity():
    """simulate -> NeuralDemixer -> Model.fit, all three on the device (README.md:28-51 of the reference)."""
    import os
    import torch
    from circuitmap_b200 import Model, NeuralDemixer, optimise
    from circuitmap_b200.simulation import simulate_batch
    from tests.conftest import GOLDEN
    res = simulate_batch([21, 22], N=100, trials=2000, H=10, connection_prob=0.1)
    assert int(res["status"].sum().item()) == 0
    dem = NeuralDemixer(path=os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz"))
    for b in range(2):
        d = dem(res["psc"][b], verbose=False)
        stim = res["powers"][res["codes"][b].long().cpu().numpy() - 1] * (res["codes"][b].cpu().numpy() > 0)
        m = Model(100)
        m.fit(d, stim, method="caviar", fit_options=dict(iters=50, seed=1, msrmp=0.4))
        truth = set(np.nonzero(res["weights"][b].cpu().numpy())[0]); got = set(np.nonzero(m.state["mu"])[0])
        assert len(got - truth) <= 1 and len(truth & got) >= 7, (sorted(truth), sorted(got))
        strong = [n for n in truth if res["weights"][b][n] >= 20]
        assert set(strong) <= got
        for n in strong:                                                 # strong weights are recovered to ~10 %
            assert abs(m.state["mu"][n] / float(res["weights"][b][n]) - 1) < 0.2
    # the uint8 codes feed the fit directly, identically to the float design
    N, f64 = 100, dict(dtype=torch.float64, device="cuda")
    cov = torch.zeros(1, N, 2, 2, **f64); cov[..., 0, 0] = 0.1; cov[..., 1, 1] = 1.0
    phi = torch.stack([0.1 * torch.ones(1, N, **f64), 5 * torch.ones(1, N, **f64)], -1).contiguous()
    pri = (torch.zeros(1, N, **f64), 10 * torch.ones(1, N, **f64), 1.0, 0.1, phi, cov)
    a = optimise.caviar_batched(res["codes"][1:2], res["powers"], *pri, psc=d[None], seeds=[1], iters=50, msrmp=0.4)
    assert np.array_equal(a["mu"][0].cpu().numpy(), m.state["mu"])


def test_unsupported_options_are_loud():
    from circuitmap_b200 import simulate
    with pytest.raises(NotImplementedError):
        simulate(N=20, trials=40, H=4, design="random")
    with pytest.raises(NotImplementedError):
        simulate(N=20, trials=40, H=4, nreps=2)
    with pytest.raises(TypeError):
        simulate(N=20, trials=40, H=4, no_such_option=1)
    with pytest.raises(RuntimeError, match="power"):
        simulate(N=20, trials=40, H=4, powers=[50, 150])
