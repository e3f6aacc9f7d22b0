"""Synthetic mapping experiments (oracle-side data generator; test infrastructure only).

`simulate` restates circuitmap/simulation.py:25-195 (blockwise/random designs) with the same
`np.random` call order; its JAX helpers are deterministic and are restated
in NumPy (simulation.py:17-23 kernel_conv_trialwise, :285-289 get_psc_kernel).
`simulate_fast` draws from the same distributions with a vectorised Generator for the large
benchmark configurations, where the reference's python double loops are infeasible.
"""
import itertools
import numpy as np


def _psc_kernels(tau_r, tau_d, T, eps=1e-5):
    """simulation.py:285-289: unit-area (trapz over T samples) bi-exponential kernels."""
    kr = np.arange(T)[None, :]
    ke = np.exp(-kr / tau_d[:, None]) - np.exp(-kr / tau_r[:, None])
    tz = np.sum(ke, 1) - 0.5 * (ke[:, 0] + ke[:, -1])
    return ke / (tz[:, None] + eps)


def _spont_kernel(tau_r, tau_d, t, delta):
    """simulation.py:200-203."""
    with np.errstate(over="ignore", invalid="ignore"):
        return (np.exp(-(t - delta) / tau_d) - np.exp(-(t - delta) / tau_r)) * (t > delta)


def _gp_cov(T, lengthscale):
    D = np.arange(T)[None, :] - np.arange(T)[:, None]
    return np.exp(-D ** 2 / (2 * lengthscale ** 2))


def _evoked(psc_kernels, spk_times, spks, mult_noise, weights, T):
    """Sum over neurons of kernel shifted to int(spike time), normalised by (sum+1e-5), times
    mult_noise * weight (simulation.py:17-23,138-155).  Only weighted, spiking pairs contribute."""
    N, K = spks.shape
    psc = np.zeros((K, T))
    for n in np.nonzero(weights)[0]:
        for k in np.nonzero(spks[n])[0]:
            s = int(spk_times[n, k])
            if s >= T:
                continue
            ke = np.zeros(T)
            ke[s:] = psc_kernels[n, :T - s]
            psc[k] += ke / (np.sum(ke) + 1e-5) * mult_noise[n, k] * weights[n]
    return psc


def simulate(N=300, T=900, H=10, trials=1000, nreps=1, connection_prob=0.05, powers=(45, 55, 65), min_latency=160,
             gamma_beta=1.5e1, sigma=6e-4, frac_strongly_connected=0.2, strong_weight_lower=20,
             strong_weight_upper=40, weak_exp_mean=4, min_weight=9, phi_0_lower=0.2, phi_0_upper=0.25,
             phi_1_lower=10, phi_1_upper=15, mult_noise_log_var=0.01, tau_r_min=25, tau_r_max=60, tau_delta_min=75,
             tau_delta_max=250, gp_scale=4e-3, gp_lengthscale=50, spont_prob=0.05, design="blockwise",
             max_power_min_spike_rate=0.4):
    """simulation.py:25-195 (weights/kernel/phi arguments omitted: the defaults are always sampled)."""
    assert design in ["random", "blockwise"]
    Trange = np.arange(T)
    if design == "blockwise":
        stim_matrix = []
        K = 0
        powers = np.sort(powers)[::-1]
        while K < trials:
            neuron_order = np.random.choice(N, N, replace=False)
            holos = [neuron_order[i * H:(i + 1) * H] for i in range(int(np.ceil(N / H)))]
            for (power, holo, rep) in itertools.product(powers, holos, range(nreps)):
                if K >= trials:
                    break
                stim_trial = np.zeros(N)
                stim_trial[holo] = power
                stim_matrix += [stim_trial]
                K += 1
        reorder = np.random.choice(K, K, replace=False)
        stim_matrix = np.array(stim_matrix).T
        stim_matrix = stim_matrix[:, reorder]
    else:
        K = trials
        stim_matrix = np.zeros((N, K))
        power_order = np.random.choice(
            np.concatenate(np.array([p * arr for p, arr in zip(powers, np.split(np.ones(K), len(powers)))])),
            K, replace=False)
        for k in range(K):
            tars = np.random.choice(N, H)
            stim_matrix[tars, k] = power_order[k]
    I = np.array([np.unique(stim_matrix[:, k])[-1] for k in range(K)])

    tau_r = np.random.uniform(tau_r_min, tau_r_max, N)
    tau_delta = np.random.uniform(tau_delta_min, tau_delta_max, N)
    tau_d = tau_r + tau_delta
    psc_kernels = _psc_kernels(tau_r, tau_d, T)

    phi_0 = np.random.uniform(phi_0_lower, phi_0_upper, N)
    phi_1 = np.random.uniform(phi_1_lower, phi_1_upper, N)
    sig = lambda x: 1 / (1 + np.exp(-x))
    frates = np.array([sig(phi_0 * stim_matrix[:, k] - phi_1) for k in range(K)]).T * (stim_matrix > 0)
    spks = (np.random.rand(N, K) <= frates).astype(float)
    noise = np.random.normal(0, sigma, [K, T])
    mult_noise = np.random.lognormal(0, mult_noise_log_var, [N, K])

    max_power = np.max(powers)
    for n in range(N):
        locs = np.where(stim_matrix[n] == max_power)[0]
        fr = np.mean(spks[n, locs])
        fr_diff = max_power_min_spike_rate - fr
        if fr_diff > 0:
            zero_locs = np.where(spks[n, locs] == 0)[0]
            req_spks = int(np.ceil(fr_diff * locs.shape[0]))
            spks[n, locs[np.random.choice(zero_locs, req_spks, replace=False)]] = 1.0

    spk_times = np.zeros((N, K))
    for n in range(N):
        for k in np.nonzero(spks[n])[0]:
            spk_times[n, k] = min_latency + np.random.gamma(1e4 / stim_matrix[n, k] ** 2, gamma_beta)

    n_connected = int(connection_prob * N)
    connected = np.random.choice(np.arange(N), n_connected, replace=False)
    n_strong = int(np.ceil(frac_strongly_connected * n_connected))
    strongly = np.random.choice(connected, n_strong, replace=False)
    weakly = np.setdiff1d(connected, strongly)
    weights = np.zeros(N)
    weights[strongly] = np.random.uniform(strong_weight_lower, strong_weight_upper, n_strong)
    weights[weakly] = np.random.exponential(weak_exp_mean, len(weakly)) + min_weight

    psc = _evoked(psc_kernels, spk_times, spks, mult_noise, weights, T)

    spont_pscs = np.zeros((K, T))
    for k in range(K):
        if np.random.rand() <= spont_prob:
            tr = np.random.uniform(tau_r_min, tau_r_max)
            tdl = np.random.uniform(tau_delta_min, tau_delta_max)
            st = np.random.randint(1, T)
            wt = np.random.uniform(np.min(weights[connected]), np.max(weights[connected]))
            kern = _spont_kernel(tr, tr + tdl, Trange, st)
            tz = np.sum(kern) - 0.5 * (kern[0] + kern[-1])
            spont_pscs[k] = wt * kern / (tz + 1e-5)

    gp_noise = gp_scale * np.random.multivariate_normal(np.zeros(T), _gp_cov(T, gp_lengthscale), size=K)
    psc = psc + spont_pscs + gp_noise + noise
    return dict(weights=weights, phi_0=phi_0, phi_1=phi_1, mult_noise=mult_noise, sigma=sigma,
                stim_matrix=stim_matrix, psc=psc, gp_noise=gp_noise, spks=spks, spk_times=spk_times,
                spont_pscs=spont_pscs, I=I)


def simulate_fast(N=1000, K=10000, T=900, H=10, connection_prob=0.1, powers=(45, 55, 65), seed=0,
                  sigma=6e-4, gp_scale=4e-3, gp_lengthscale=50, spont_prob=0.05, mult_noise_log_var=0.01,
                  gamma_beta=1.5e1, min_latency=160, dtype=np.float64):
    """Vectorised generator with simulate()'s distributions (blockwise design, nreps=1); NOT
    stream-compatible with it.  Used for the benchmark-sized configurations."""
    rng = np.random.default_rng(seed)
    powers = np.sort(np.asarray(powers, dtype=np.float64))[::-1]
    nh = int(np.ceil(N / H))
    cols_n, cols_p = [], []
    Kc = 0
    while Kc < K:
        order = rng.permutation(N)
        for p in powers:
            for h in range(nh):
                if Kc >= K:
                    break
                cols_n.append(order[h * H:(h + 1) * H])
                cols_p.append(p)
                Kc += 1
    reorder = rng.permutation(K)
    stim = np.zeros((N, K), dtype=np.float64)
    for j, k in enumerate(reorder):
        stim[cols_n[k], j] = cols_p[k]

    tau_r = rng.uniform(25, 60, N)
    tau_d = tau_r + rng.uniform(75, 250, N)
    kern = _psc_kernels(tau_r, tau_d, T)
    phi_0 = rng.uniform(0.2, 0.25, N)
    phi_1 = rng.uniform(10, 15, N)
    n_conn = int(connection_prob * N)
    connected = rng.choice(N, n_conn, replace=False)
    n_strong = int(np.ceil(0.2 * n_conn))
    weights = np.zeros(N)
    weights[connected[:n_strong]] = rng.uniform(20, 40, n_strong)
    weights[connected[n_strong:]] = rng.exponential(4, n_conn - n_strong) + 9

    psc = np.zeros((K, T))
    spks = np.zeros((N, K), dtype=bool)
    for n in range(N):
        ks = np.nonzero(stim[n])[0]
        fr = 1 / (1 + np.exp(-(phi_0[n] * stim[n, ks] - phi_1[n])))
        sp = rng.random(ks.size) <= fr
        mx = ks[stim[n, ks] == powers[0]]
        if mx.size:
            have = sp[stim[n, ks] == powers[0]]
            need = int(np.ceil(max(0.0, 0.4 - have.mean()) * mx.size))
            if need > 0:
                zl = np.nonzero(~have)[0]
                pick = rng.choice(zl, need, replace=False)
                idx = np.nonzero(stim[n, ks] == powers[0])[0][pick]
                sp[idx] = True
        spks[n, ks] = sp
        if weights[n] == 0:
            continue
        for k in ks[sp]:
            s = int(min_latency + rng.gamma(1e4 / stim[n, k] ** 2, gamma_beta))
            if s >= T:
                continue
            ke = np.zeros(T)
            ke[s:] = kern[n, :T - s]
            psc[k] += ke / (ke.sum() + 1e-5) * rng.lognormal(0, mult_noise_log_var) * weights[n]
    Trange = np.arange(T)
    for k in np.nonzero(rng.random(K) <= spont_prob)[0]:
        tr = rng.uniform(25, 60)
        td = tr + rng.uniform(75, 250)
        st = rng.integers(1, T)
        wt = rng.uniform(weights[connected].min(), weights[connected].max())
        kk = _spont_kernel(tr, td, Trange, st)
        psc[k] += wt * kk / (kk.sum() - 0.5 * (kk[0] + kk[-1]) + 1e-5)
    Lc = np.linalg.cholesky(_gp_cov(T, gp_lengthscale) + 1e-6 * np.eye(T))
    psc += gp_scale * (rng.standard_normal((K, T)) @ Lc.T)
    psc += rng.normal(0, sigma, (K, T))
    return dict(psc=psc.astype(dtype), stim_matrix=stim.astype(dtype), weights=weights, spks=spks,
                phi_0=phi_0, phi_1=phi_1)
