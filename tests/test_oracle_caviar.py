"""Oracle self-checks: literal (reference structure) == reduced (kernel structure); PAVA; invariants; regression pin."""
import os

import numpy as np
import pytest

from oracle import caviar as oc, simulate as osim
from oracle.pava import isotonic_regression, pava_last
from tests.conftest import GOLDEN


def test_pava_matches_sklearn_and_last_pool():
    from sklearn.isotonic import IsotonicRegression
    rng = np.random.default_rng(0)
    for _ in range(50):
        y = rng.random(rng.integers(1, 8))
        fit = isotonic_regression(y)
        want = IsotonicRegression().fit_transform(np.arange(len(y)), y)
        assert np.allclose(fit, want, atol=1e-12)
        assert abs(pava_last(y) - fit[-1]) < 1e-15
    assert isotonic_regression([0.3, 0.2, 0.1]).tolist() == pytest.approx([0.2, 0.2, 0.2])
    assert pava_last([0.0, 0.0, 0.0]) == 0.0


@pytest.fixture(scope="module")
def tiny():
    np.random.seed(3)
    return osim.simulate(N=32, trials=300, H=4, connection_prob=0.15)


def _run(sim, form, trace=None, **kw):
    pr = oc.default_priors(sim["stim_matrix"].shape[0])
    opts = dict(iters=12, seed=1, msrmp=0.4)
    opts.update(kw)
    return oc.caviar(sim["psc"], sim["stim_matrix"], pr["mu"], pr["beta"], pr["shape"], pr["rate"], pr["phi"],
                     pr["phi_cov"], form=form, trace=trace, **opts)


def test_literal_equals_reduced(tiny):
    t1, t2 = {"decisions": []}, {"decisions": []}
    a = _run(tiny, "literal", t1)
    b = _run(tiny, "reduced", t2)
    for i in range(8):
        x, y = np.asarray(a[i], float), np.asarray(b[i], float)
        assert np.allclose(x, y, rtol=1e-7, atol=1e-9), i
    assert [(d[2], d[6]) for d in t1["decisions"]] == [(d[2], d[6]) for d in t2["decisions"]]
    assert np.array_equal(a[0] != 0, b[0] != 0)


def test_invariants(tiny):
    K = tiny["psc"].shape[0]
    tr = {"decisions": []}
    res = _run(tiny, "reduced", tr, iters=24, save_histories=True)
    mu, beta, lam, shape, rate = res[:5]
    I = tiny["stim_matrix"]
    assert np.all(lam[I == 0] == 0)                                  # supp(lam) within supp(I)
    assert shape == 1.0 + K / 2                                      # caviar.py:241
    z_hist = res[16]
    assert np.array_equal(z_hist[20], z_hist[23])                    # soft-threshold frozen for it >= 20 (A.3 #5)
    lam_hist, mu_hist = res[11], res[9]
    for it in range(3, 24):                                          # mu[n]==0 <=> lam row of previous iteration == 0
        assert np.array_equal(mu_hist[it] == 0, lam_hist[it - 1].sum(1) == 0)
    assert res[8] is None and res[12].shape == (24, K)


def test_regression_pin():
    g = np.load(os.path.join(GOLDEN, "caviar_oracle_pin_N32_K300.npz"))
    np.random.seed(3)
    sim = osim.simulate(N=32, trials=300, H=4, connection_prob=0.15)
    st = oc.fit(sim["psc"], sim["stim_matrix"], iters=30, seed=1, msrmp=0.4)
    for k in ["mu", "beta", "shape", "rate", "phi", "phi_cov", "z"]:
        assert np.allclose(st[k], g[k], rtol=1e-9, atol=1e-12), k
    assert np.array_equal(np.nonzero(st["mu"])[0], np.nonzero(sim["weights"])[0])   # recovers the true connections


def test_simulate_shapes_and_design():
    np.random.seed(0)
    sim = osim.simulate(N=40, trials=120, H=10, connection_prob=0.1)
    stim, psc = sim["stim_matrix"], sim["psc"]
    assert stim.shape == (40, 120) and psc.shape == (120, 900)
    assert not stim.flags.c_contiguous                                # as the reference returns it (simulation.py:62-63)
    assert set(np.unique(stim)) == {0.0, 45.0, 55.0, 65.0}
    assert np.all((stim > 0).sum(0) == 10)
    f = osim.simulate_fast(N=40, K=120, H=10, seed=1)
    assert f["stim_matrix"].shape == (40, 120) and np.all((f["stim_matrix"] > 0).sum(0) == 10)


def test_literal_equals_reduced_for_large_powers():
    """for powers >= ~150 the reference's log(f / (1 - f)) Monte-Carlo term saturates (f rounds to 1 -> +inf) and no
    longer equals the linear form; the reduced form (= the kernel) switches to the reference's expression there."""
    sim = osim.simulate_fast(N=24, K=240, H=4, seed=5, powers=(60, 90, 99))
    stim = sim["stim_matrix"].copy()
    stim[stim == 90] = 160.0                       # the same experiment reported with larger power values
    stim[stim == 99] = 320.0
    t1, t2 = {"decisions": []}, {"decisions": []}
    a = _run(dict(sim, stim_matrix=stim), "literal", t1, iters=8)
    b = _run(dict(sim, stim_matrix=stim), "reduced", t2, iters=8)
    assert [(d[2], d[6]) for d in t1["decisions"]] == [(d[2], d[6]) for d in t2["decisions"]]
    for i in range(8):
        x, y = np.asarray(a[i], float), np.asarray(b[i], float)
        assert np.allclose(x, y, rtol=1e-6, atol=1e-9, equal_nan=True), i
    mc = np.c_[np.full(100, 0.3), np.full(100, 5.0)]
    out = oc.mc_term_per_power(mc, np.array([50.0, 160.0]))
    assert np.isclose(out[0], 10.0) and np.isinf(out[1])          # 0.3 * 160 - 5 = 43 > 36.7: f == 1 in float64
