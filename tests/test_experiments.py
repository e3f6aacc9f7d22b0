"""Batched experiment drivers (circuitmap_b200/experiments.py) against the reference scripts' own expressions (CPU)
and against looping the single-fit API (GPU)."""
import numpy as np
import pytest


def test_trial_count_schedule_matches_script_expression():
    from circuitmap_b200.experiments import downsample_trial_counts
    for K, dstime in ((2000, 10), (1999, 7), (300, 10), (250, 60)):
        ds_step = dstime * 30                                                   # run_downsampling_experiments.py:68-72
        want = np.concatenate([np.arange(ds_step, K + 1, ds_step), [K]])
        assert np.array_equal(downsample_trial_counts(K, dstime), want)


def test_unique_holograms_matches_script_expression():
    from circuitmap_b200.experiments import unique_holograms
    rng = np.random.default_rng(0)
    N, K = 12, 90
    stim = np.zeros((N, K))
    holos = [rng.choice(N, 3, replace=False) for _ in range(7)]
    for k in range(K):
        if k % 10 == 0:
            stim[rng.integers(N), k] = 45.0                                     # single-target trials are ignored
        else:
            stim[holos[rng.integers(7)], k] = rng.choice([45.0, 55.0, 65.0])
    uniq, ids, multi = unique_holograms(stim)
    stim_bin = (stim[:, np.where(np.sum(stim > 0, axis=0) > 1)[0]] != 0).astype(float)
    want = np.vstack(list({tuple(row) for row in stim_bin.T}))                  # generate_loho_cv_slurm_scripts.py:110-112
    assert len(uniq) == len(want) and {tuple(r) for r in uniq} == {tuple(r) for r in want}
    assert np.array_equal(uniq[ids], stim_bin.T) and multi.size == stim_bin.shape[1]


def test_result_file_name_and_key(tmp_path):
    from circuitmap_b200.experiments import save_downsampling
    w = np.arange(24.0).reshape(2, 3, 4)
    path = save_downsampling(str(tmp_path), "/data/exp_0420_cell3.mat", 10, 2, "multi", "caviar", w)
    assert path.endswith("exp_0420_cell3_downsampling_weights_steptime10_nreps2_designmulti_methodcaviar.npz")
    assert np.array_equal(np.load(path)["weights"], w)


@pytest.mark.gpu
def test_downsampling_and_loho_equal_looping_model_fit():
    from circuitmap_b200 import Model
    from circuitmap_b200.experiments import downsampling_weights, loho_cv_weights, unique_holograms
    from oracle import simulate as osim
    sim = osim.simulate_fast(N=30, K=900, H=4, seed=2)
    stim, psc = np.ascontiguousarray(sim["stim_matrix"]), sim["psc"]
    N, K = stim.shape
    got = downsampling_weights(psc, stim, dstime=10, n_repeats=2, msrmp=0.4, rng=np.random.RandomState(5))
    rng = np.random.RandomState(5)
    counts = np.concatenate([np.arange(300, K + 1, 300), [K]])
    assert got.shape == (2, len(counts), N)
    for st, cnt in enumerate(counts):                                           # the script's loop, one Model per fit
        for r in range(2):
            trials = rng.choice(K, cnt, replace=False)
            m = Model(N)
            m.fit(psc[trials], stim[:, trials], method="caviar",
                  fit_options={"save_histories": False, "tol": 0.005, "msrmp": 0.4, "fn_scan": True})
            assert np.array_equal(got[r, st], m.state["mu"]), (st, r)
    uniq, ids, multi = unique_holograms(stim)
    folds = [0, 3, len(uniq) - 1]
    mu, f = loho_cv_weights(psc, stim, msrmp=0.4, hologram_ids=folds)
    for row, h in zip(mu, folds):
        keep = multi[ids != h]
        m = Model(N)
        m.fit(psc[keep], stim[:, keep], method="caviar", fit_options={"save_histories": False, "msrmp": 0.4})
        assert np.array_equal(row, m.state["mu"])


def test_main_script_writers(tmp_path):
    """run_circuitmap_main.py:50-63: file names, keys and the .mat round trip."""
    from scipy.io import loadmat
    from circuitmap_b200.experiments import load_experiment, save_main_results
    state = {"mu": np.arange(5.0), "beta": np.ones(5), "lam": np.eye(5, 7)}
    mat, npz = save_main_results(str(tmp_path), "/data/cell12_map.npz", state)
    assert mat.endswith("cell12_map_cmap.mat") and npz.endswith("cell12_map_cmap.npz")
    a, b = loadmat(mat), np.load(npz)
    for k, v in (("weights", state["mu"]), ("weight_uncertainty", state["beta"]), ("spikes", state["lam"])):
        assert np.array_equal(np.squeeze(a[k]), v) and np.array_equal(b[k], v)
    np.savez(str(tmp_path / "exp.npz"), psc=np.zeros((4, 900)), stimulus_matrix=np.zeros((3, 4)))
    psc, stim = load_experiment(str(tmp_path / "exp.npz"))
    assert psc.shape == (4, 900) and stim.shape == (3, 4)
    with pytest.raises(Exception):
        load_experiment(str(tmp_path / "exp.csv"))


@pytest.mark.gpu
def test_downsampling_subset_and_loho_fold_against_the_oracle():
    """One downsampling subset and one leave-one-hologram-out fold of the batched drivers, checked against the CPU oracle
    (not against the CUDA path itself)."""
    from circuitmap_b200.experiments import downsampling_weights, loho_cv_weights, unique_holograms
    from oracle import caviar as oc, simulate as osim
    sim = osim.simulate_fast(N=30, K=900, H=4, seed=2)
    stim, psc = np.ascontiguousarray(sim["stim_matrix"]), sim["psc"]
    N, K = stim.shape
    got = downsampling_weights(psc, stim, dstime=10, n_repeats=2, msrmp=0.4, rng=np.random.RandomState(5))
    rng = np.random.RandomState(5)
    draws = [[rng.choice(K, cnt, replace=False) for _ in range(2)] for cnt in (300, 600, 900, 900)]
    for st, r in ((0, 1), (2, 0)):
        tr = draws[st][r]
        ref = oc.fit(psc[tr], stim[:, tr], save_histories=False, tol=0.005, msrmp=0.4, fn_scan=True)
        assert np.array_equal(got[r, st] != 0, ref["mu"] != 0)
        assert np.allclose(got[r, st], ref["mu"], rtol=1e-4, atol=1e-7 * np.abs(ref["mu"]).max())
    uniq, ids, multi = unique_holograms(stim)
    mu, _ = loho_cv_weights(psc, stim, msrmp=0.4, hologram_ids=[3])
    keep = multi[ids != 3]
    ref = oc.fit(psc[keep], stim[:, keep], save_histories=False, msrmp=0.4)
    assert np.array_equal(mu[0] != 0, ref["mu"] != 0)
    assert np.allclose(mu[0], ref["mu"], rtol=1e-4, atol=1e-7 * np.abs(ref["mu"]).max())


@pytest.mark.gpu
def test_run_main_end_to_end(tmp_path):
    """scripts/run_circuitmap_main.py as a function: .npz in, demix -> fit on the device, .mat / .npz out; equals the two
    drop-in calls made by hand with NumPy arrays."""
    import os
    from scipy.io import loadmat
    from circuitmap_b200 import Model, NeuralDemixer
    from circuitmap_b200.experiments import run_main
    from oracle import simulate as osim
    from tests.conftest import GOLDEN
    sim = osim.simulate_fast(N=30, K=400, H=4, seed=6)
    data = str(tmp_path / "exp7.npz")
    np.savez(data, psc=sim["psc"], stimulus_matrix=sim["stim_matrix"])
    wpath = os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz")
    model, (mat, npz) = run_main(data, {"demixer": wpath, "msrmp": 0.4}, str(tmp_path))
    dem = NeuralDemixer(path=wpath)(sim["psc"], verbose=False)
    m2 = Model(30)
    m2.fit(dem, sim["stim_matrix"], method="caviar", fit_options={"msrmp": 0.4, "save_histories": False})
    assert np.array_equal(model.state["mu"] != 0, m2.state["mu"] != 0)
    for k in ("mu", "beta", "lam"):
        assert np.allclose(model.state[k], m2.state[k], rtol=1e-9, atol=1e-12)
    out = loadmat(mat)
    assert np.array_equal(np.squeeze(out["weights"]), model.state["mu"]) and out["spikes"].shape == (30, 400)
    assert np.array_equal(np.load(npz)["spikes"], model.state["lam"])
