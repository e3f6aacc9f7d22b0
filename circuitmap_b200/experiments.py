"""Batched drivers for the reference's experiment scripts.

The reference runs these as python loops / SLURM arrays of independent `Model.fit` calls; here every group of fits
with the same trial count is ONE `cm_caviar_fit` call (B fits), so a sweep costs a few launches.

  downsampling_weights   scripts/run_downsampling_experiments.py:66-91  (trial-count schedule :68-72, subset draw :78,
                         fit options :83, result file :93-97)
  unique_holograms       scripts/generate_loho_cv_slurm_scripts.py:108-112 (leave-one-hologram-out fold enumeration)
  load_experiment / save_main_results / run_main   scripts/run_circuitmap_main.py:21-63 (NeuroCAAS entry: reader, demix ->
                         fit, .mat / .npz writers)
  loho_cv_weights        the per-fold worker the reference's SLURM script calls (`run_loho_cv_caviar.py`) is NOT in the
                         reference repository; the fold definition below (drop every trial whose binarised hologram
                         equals the held-out one, fit on the rest) follows the script generator's enumeration.
"""
import os

import numpy as np

from . import _staging, optimise


def downsample_trial_counts(K, dstime, stim_freq=30):
    """Numbers of trials of the downsampling steps (run_downsampling_experiments.py:68-72)."""
    ds_step = dstime * stim_freq
    return np.concatenate([np.arange(ds_step, K + 1, ds_step), [K]])


def _default_priors(N, B, dev):
    import torch
    f64 = dict(dtype=torch.float64, device=dev)
    cov = torch.zeros(B, N, 2, 2, **f64)
    cov[..., 0, 0] = 1e-1
    cov[..., 1, 1] = 1e0
    phi = torch.stack([1e-1 * torch.ones(B, N, **f64), 5e0 * torch.ones(B, N, **f64)], -1).contiguous()
    return torch.zeros(B, N, **f64), 1e1 * torch.ones(B, N, **f64), 1.0, 1e-1, phi, cov      # model.py:24-31


def _fit_subsets(psc_dev, stim_dev, subsets, powers, fit_options, seed):
    """Batched calls for trial subsets of equal length -> mu (B, N) as a NumPy array.  The reference derives the power
    levels from each subset (`np.unique(I)[1:]`, caviar.py:42): subsets that miss a level are fitted with their own."""
    import torch
    N = stim_dev.shape[0]
    idx = torch.as_tensor(np.stack(subsets), device=stim_dev.device)                          # (B, Ks)
    stim_b = stim_dev[:, idx].permute(1, 0, 2).contiguous()                                   # (B, N, Ks)
    psc_b = psc_dev[idx].contiguous()                                                         # (B, Ks, T)
    present = torch.stack([(stim_b == float(p)).flatten(1).any(1) for p in powers], 1).cpu().numpy()   # (B, P)
    mu = np.zeros((len(subsets), N))
    groups = {}
    for b, row in enumerate(present):
        groups.setdefault(tuple(row), []).append(b)
    for mask, members in groups.items():
        sel = torch.as_tensor(members, device=stim_dev.device)
        out = optimise.caviar_batched(stim_b[sel].contiguous(), np.asarray(powers)[list(mask)],
                                      *_default_priors(N, len(members), stim_dev.device), psc=psc_b[sel].contiguous(),
                                      seeds=[seed] * len(members), want_lam=False, **fit_options)
        optimise.check_status(out)
        mu[members] = out["mu"].cpu().numpy()
    return mu


def downsampling_weights(psc_dem, stim_matrix, dstime, n_repeats, msrmp, stim_freq=30, rng=np.random, device=None,
                         seed=0):
    """The loop of run_downsampling_experiments.py:74-91 for method='caviar': for every trial count of the schedule and
    every repeat, draw `trials = rng.choice(K, count, replace=False)` (same draw order as the script: steps outer,
    repeats inner) and fit a fresh `Model(N)` with fit_options {'save_histories': False, 'tol': 0.005, 'msrmp': msrmp,
    'fn_scan': True}.  Returns estimated_weights (n_repeats, nsteps, N) = model.state['mu'] of every fit."""
    import torch
    dev = torch.device("cuda" if device is None else device)
    stim_matrix = np.ascontiguousarray(stim_matrix, dtype=np.float64)
    psc_dem = np.ascontiguousarray(psc_dem)
    N, K = stim_matrix.shape
    counts = downsample_trial_counts(K, dstime, stim_freq)
    powers = np.unique(stim_matrix)[1:]
    with torch.cuda.device(dev):
        stim_dev = _staging.to_device(stim_matrix, dev)
        psc_dev = _staging.to_device(psc_dem, dev)
    opts = {"save_histories": False, "tol": 0.005, "msrmp": msrmp, "fn_scan": True}
    weights = np.zeros((n_repeats, len(counts), N))
    for st, cnt in enumerate(counts):
        subsets = [rng.choice(K, int(cnt), replace=False) for _ in range(n_repeats)]
        weights[:, st] = _fit_subsets(psc_dev, stim_dev, subsets, powers, opts, seed)
    return weights


def save_downsampling(out, data_path, dstime, n_repeats, design, method, weights):
    """Result file of run_downsampling_experiments.py:93-97 (same name pattern and key)."""
    if out[-1] != "/":
        out += "/"
    base = os.path.basename(data_path)[:-4]
    path = out + base + "_downsampling_weights_steptime%i_nreps%i_design%s_method%s" % (dstime, n_repeats, design, method)
    np.savez(path, weights=weights)
    return path + ".npz"


def unique_holograms(stim_matrix):
    """Binarised multi-target holograms of an experiment (generate_loho_cv_slurm_scripts.py:104-112): returns
    (unique (H, N) array in first-appearance order, hologram id of every multi-target trial, indices of those trials)."""
    stim_matrix = np.asarray(stim_matrix, dtype=float)
    multi = np.where(np.sum(stim_matrix > 0, axis=0) > 1)[0]
    stim_bin = (stim_matrix[:, multi] != 0).astype(float)
    seen, ids, uniq = {}, np.zeros(multi.size, dtype=int), []
    for k, row in enumerate(stim_bin.T):
        key = row.tobytes()
        if key not in seen:
            seen[key] = len(uniq)
            uniq.append(row)
        ids[k] = seen[key]
    return np.array(uniq), ids, multi


def loho_cv_weights(psc_dem, stim_matrix, msrmp, hologram_ids=None, device=None, seed=0, max_batch=64):
    """Leave-one-hologram-out fits: fold h trains on every multi-target trial whose hologram is not h.  Folds with the
    same number of training trials share one batched call.  Returns (mu (n_folds, N), fold hologram ids)."""
    import torch
    dev = torch.device("cuda" if device is None else device)
    stim_matrix = np.ascontiguousarray(stim_matrix, dtype=np.float64)
    uniq, ids, multi = unique_holograms(stim_matrix)
    folds = np.arange(len(uniq)) if hologram_ids is None else np.asarray(hologram_ids, dtype=int)
    stim_multi = np.ascontiguousarray(stim_matrix[:, multi])
    powers = np.unique(stim_multi)[1:]
    with torch.cuda.device(dev):
        stim_dev = _staging.to_device(stim_multi, dev)
        psc_dev = _staging.to_device(np.ascontiguousarray(np.asarray(psc_dem)[multi]), dev)
    opts = {"save_histories": False, "msrmp": msrmp}
    N = stim_matrix.shape[0]
    mu = np.zeros((len(folds), N))
    train = {int(h): np.nonzero(ids != h)[0] for h in folds}
    by_len = {}
    for pos, h in enumerate(folds):
        by_len.setdefault(train[int(h)].size, []).append(pos)
    for _, poss in sorted(by_len.items()):
        for b0 in range(0, len(poss), max_batch):
            chunk = poss[b0:b0 + max_batch]
            mu[chunk] = _fit_subsets(psc_dev, stim_dev, [train[int(folds[p])] for p in chunk], powers, opts, seed)
    return mu, folds


# ------------------------------------------------------------------------------------------------ run_circuitmap_main.py
def load_experiment(path):
    """Input reader of scripts/run_circuitmap_main.py:21-32: .mat (scipy loadmat) or .npy / .npz (np.load) holding
    'psc' (K, T) and 'stimulus_matrix' (N, K); any other extension raises, as the script does."""
    ext = path[-4:]
    if ext == ".mat":
        from scipy.io import loadmat
        f = loadmat(path)
    elif ext in (".npy", ".npz"):
        f = np.load(path)
    else:
        raise Exception
    return f["psc"], f["stimulus_matrix"]


def save_main_results(out, data_path, state):
    """Result writers of scripts/run_circuitmap_main.py:50-63: `<out>/<stem>_cmap.mat` and `<out>/<stem>_cmap.npz`
    with keys weights = mu, weight_uncertainty = beta, spikes = lam.  Returns both paths."""
    from pathlib import Path
    from scipy.io import savemat
    if out[-1] != "/":
        out += "/"
    save_name = out + Path(data_path).stem + "_cmap"
    payload = {"weights": np.asarray(state["mu"]), "weight_uncertainty": np.asarray(state["beta"]),
               "spikes": np.asarray(state["lam"])}
    savemat(save_name + ".mat", payload)
    np.savez(save_name, **payload)
    return save_name + ".mat", save_name + ".npz"


def run_main(data, config, out, device=None):
    """scripts/run_circuitmap_main.py as a function: read the experiment, demix with the configured network, fit with the
    configured msrmp (fit_options {'msrmp': msrmp, 'save_histories': False}), write the .mat and .npz result files.
    `config` is the yaml path the script takes (keys 'demixer', 'msrmp') or an equivalent dict.  The demixed traces stay
    on the device between the two steps."""
    import torch
    from . import Model, NeuralDemixer
    if not isinstance(config, dict):
        import yaml
        config = yaml.safe_load(open(config))
    psc, stim_matrix = load_experiment(data)
    demix = NeuralDemixer(path=config["demixer"], device=device)
    psc_dev = torch.from_numpy(np.ascontiguousarray(psc, dtype=np.float64)).to(demix.device)
    psc_dem = demix(psc_dev)                                    # device tensor carrying y / sum-of-squares for the fit
    msrmp = float(config["msrmp"])
    N = stim_matrix.shape[0]
    model = Model(N)
    model.fit(psc_dem, stim_matrix, method="caviar", fit_options={"msrmp": msrmp, "save_histories": False})
    return model, save_main_results(out, data, model.state)
