"""CAVIaR restated in NumPy fp64 (oracle; test infrastructure only).

Follows circuitmap/optimise/caviar.py line by line (citations on each function).  JAX is
absent here, so this is a restatement; it is PINNED against the reference's own source
executed through a NumPy-backed JAX stand-in (oracle/jax_shim.py, fixtures
tests/golden/caviar_ref_*.npz, tests/test_reference_pin.py; see oracle/__init__.py for what
that does and does not cover).  Two forms are provided and proven equal in tests:

  form='literal'  the reference's own structure: O(N^2 K) masked sum per neuron
                  (caviar.py:204-206), S x K Monte-Carlo logit(sigmoid()) term
                  (caviar.py:215,233-235), K-term Newton sums (caviar.py:275-287,314-316).
  form='reduced'  algebraically identical O(nnz) form: running
                  prediction vector, mean-of-samples MC term, per-power sufficient
                  statistics for the Newton step.  This is the form the CUDA kernel
                  implements and the fair CPU baseline.

All load-bearing quirks of the reference are kept (variance used as sd, beta squared
twice, in-sweep mu zeroing discarded, soft-threshold loop counter starting at `it`, ...).
"""
import numpy as np
from scipy.special import ndtr, ndtri
from scipy.stats import sem

from . import prng
from .pava import isotonic_regression, pava_last

DBL_MAX = np.finfo(np.float64).max


def sigmoid(x):
    with np.errstate(over="ignore"):
        return 1.0 / (1.0 + np.exp(-x))


def trapz_rows(psc):
    """np.trapz(psc, axis=-1) with unit spacing (caviar.py:28)."""
    psc = np.asarray(psc, dtype=np.float64)
    return np.sum(psc, axis=-1) - 0.5 * (psc[..., 0] + psc[..., -1])


def prologue(y_psc, I, y_xcorr_thresh):
    """caviar.py:28-35,42: y, lam_mask, lam0, powers."""
    y_psc = np.asarray(y_psc, dtype=np.float64)
    y = trapz_rows(y_psc)
    lam_mask = (np.sum(y_psc * y_psc, axis=-1) > y_xcorr_thresh).astype(np.float64)
    I = np.ascontiguousarray(np.asarray(I, dtype=np.float64))
    lam = np.zeros_like(I)
    lam[I > 0] = 0.95
    lam = lam * lam_mask
    powers = np.unique(I)[1:]
    return y, lam_mask, lam, I, powers


# --------------------------------------------------------------------------- a2
def block_update_mu(y, lam, shape, rate, mu_prior, beta_prior):
    """caviar.py:166-172 (dense, literal)."""
    N = lam.shape[0]
    L = lam @ lam.T
    D = np.diag(np.sum(lam * (1 - lam), axis=-1))
    cov = np.linalg.inv(shape / rate * (D + L) + 1 / (beta_prior ** 2) * np.eye(N))
    mean = cov @ (shape / rate * np.sum(y * lam, axis=1) + 1 / (beta_prior ** 2) * mu_prior)
    return mean, np.diag(cov).copy()


def block_update_mu_active(y, lam, shape, rate, mu_prior, beta_prior):
    """Same system solved on the active set only (App. A.2): rows with lam[n,:]==0 decouple."""
    sig = shape / rate
    act = np.nonzero(np.any(lam != 0, axis=1))[0]
    mu = np.array(mu_prior, dtype=np.float64).copy()
    beta = np.asarray(beta_prior, dtype=np.float64) ** 2
    beta = beta.copy()
    if act.size:
        la = lam[act]
        M = sig * (np.diag(np.sum(la * (1 - la), axis=-1)) + la @ la.T) + np.diag(1 / beta_prior[act] ** 2)
        b = sig * (la @ y) + mu_prior[act] / beta_prior[act] ** 2
        C = np.linalg.inv(M)
        mu[act] = C @ b
        beta[act] = np.diag(C)
    return mu, beta


# --------------------------------------------------------------------------- a3-a5
def eval_spike_rates(stimv, lamv, powers):
    """caviar.py:174-186."""
    out = np.zeros(len(powers))
    for p, power in enumerate(powers):
        m = stimv == power
        cnt = np.sum(m)
        out[p] = np.sum(lamv[m]) / (cnt + 1e-4 * (cnt == 0))
    return out


def mc_samples(key, phi_n, phi_cov_n, S):
    """caviar.py:209-212: truncated-normal samples; sdev is diag(phi_cov) (a variance)."""
    u = prng.uniform_f64(key, (S, 2))
    mean, sdev = phi_n, np.diag(phi_cov_n)
    c = ndtr(-mean / sdev)
    return ndtri(c + u * (1 - c)) * sdev + mean


MC_LITERAL_X = 28.0


def mc_term_per_power(mc, powers):
    """mean_s log(f_s / (1 - f_s)), f_s = sigmoid(phi0_s I - phi1_s)  (caviar.py:213-215,233-235) for every power I.

    log(f / (1 - f)) is its argument, so the mean is mean(phi0) I - mean(phi1) -- until a sample saturates the float64
    sigmoid (the reference's value carries rounding noise 2^-53 e^x per sample: 2e-4 at x = 28; x >~ 36.7: +inf).  Reduced
    form = what the kernel does (csrc/caviar_fit.inl mc_means): linear while the largest possible argument
    max(phi0) I - min(phi1) <= 28, the reference's expression otherwise."""
    pb0, pb1 = np.mean(mc[:, 0]), np.mean(mc[:, 1])
    out = pb0 * powers - pb1
    xmax = np.max(mc[:, 0]) * powers - np.min(mc[:, 1])
    for p in np.nonzero(~(xmax <= MC_LITERAL_X))[0]:
        fn = sigmoid(mc[:, 0] * powers[p] - mc[:, 1])
        with np.errstate(divide="ignore", invalid="ignore"):
            out[p] = np.mean(np.log(fn / (1 - fn)))
    return out


def update_lam(y, I, mu, beta, lam, shape, rate, phi, phi_cov, lam_mask, key, S, powers,
               minimum_spike_count, minimax_spk_prob, it, delay_spont_est, form, trace=None):
    """caviar.py:190-231.  Returns (lam, key); in-sweep mu zeroing is discarded (:229-231)."""
    N, K = I.shape
    order = prng.permutation(key, N)
    sig = shape / rate
    lam = lam.copy()
    mu = mu.copy()
    if form == "reduced":
        pred = mu @ lam
    for m in range(N):
        n = order[m]
        ks = prng.split(key)
        key_s, key_next = ks[0], ks[1]
        mc = mc_samples(key_s, phi[n], phi_cov[n], S)
        if form == "literal":
            others = np.arange(N) != n
            arg = -2 * sig * y * mu[n] + 2 * mu[n] * np.sum(sig * mu[others, None] * lam[others], 0) \
                + sig * (mu[n] ** 2 + beta[n] ** 2)
            fn = sigmoid(mc[:, 0:1] * I[n][None, :] - mc[:, 1:2])
            with np.errstate(divide="ignore", invalid="ignore"):
                mcE = np.mean(np.log(fn / (1 - fn)), 0)
            est = lam_mask * (I[n] > 0) * sigmoid(mcE - 0.5 * arg)
        else:
            idx = np.nonzero(I[n] > 0)[0]
            arg = -2 * sig * y[idx] * mu[n] + 2 * sig * mu[n] * (pred[idx] - mu[n] * lam[n, idx]) \
                + sig * (mu[n] ** 2 + beta[n] ** 2)
            mce = mc_term_per_power(mc, powers)
            est = np.zeros(K)
            est[idx] = lam_mask[idx] * sigmoid(mce[np.searchsorted(powers, I[n, idx])] - 0.5 * arg)
        srates = eval_spike_rates(I[n], est, powers)
        pv = isotonic_regression(srates)[-1]
        tot = np.sum(est)
        ok = float(pv >= minimax_spk_prob) * float(tot >= minimum_spike_count)
        ok = ok * (it > delay_spont_est) + 1.0 * (it <= delay_spont_est)
        if trace is not None:
            trace.append((it, m, int(n), pv, tot, minimax_spk_prob, ok))
        new_row = est * ok
        if form == "reduced":
            pred = pred + (mu[n] * (ok == 1.0)) * new_row - mu[n] * lam[n]
        lam[n] = new_row
        mu[n] = mu[n] * (ok == 1.0)
        key = key_next
    return lam, key


# --------------------------------------------------------------------------- a6
def update_sigma(y, mu, beta, lam, shape_prior, rate_prior):
    """caviar.py:238-244."""
    K = y.shape[0]
    shape = shape_prior + K / 2
    rate = rate_prior + 0.5 * (np.sum(np.square(y - np.sum(mu[:, None] * lam, 0)))
                               - np.sum(np.square(mu[:, None] * lam))
                               + np.sum((mu ** 2 + beta ** 2)[:, None] * lam))
    return shape, rate


# --------------------------------------------------------------------------- a7
def _nll_literal(yv, phi, phi_prior, prec, Iv, t):
    """caviar.py:312-316 for a batch of neurons: yv (n,K), phi (n,2)."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        f = sigmoid(phi[:, 0:1] * Iv - phi[:, 1:2])
        ll = np.nan_to_num(yv * np.log(f) + (1 - yv) * np.log(1 - f))
        d = phi - phi_prior
        quad = 0.5 * np.einsum("ni,nij,nj->n", d, prec, d)
        return -np.sum(ll, axis=1) - np.sum(np.log(phi), axis=1) / t + quad


def suff_stats(lam, I, powers):
    """Per-(neuron, power group) statistics for the reduced Newton step.

    Group 0 is I==0 (untargeted trials), groups 1..P the distinct powers.
    Returns cnt, S=sum lam, n0=#(lam==0), n1=#(lam==1), each (N, P+1), and pvals (P+1,).
    """
    N, K = I.shape
    P = len(powers)
    pvals = np.concatenate([[0.0], powers])
    cnt = np.zeros((N, P + 1))
    S = np.zeros((N, P + 1))
    n0 = np.zeros((N, P + 1))
    n1 = np.zeros((N, P + 1))
    for g, pv in enumerate(pvals):
        m = I == pv
        cnt[:, g] = m.sum(1)
        S[:, g] = (lam * m).sum(1)
        n0[:, g] = ((lam == 0) & m).sum(1)
        n1[:, g] = ((lam == 1) & m).sum(1)
    return cnt, S, n0, n1, pvals


def _group_loglik(f, cnt, S, n0, n1):
    """sum_k nan_to_num(y log f + (1-y) log(1-f)) over a power group, from its statistics.

    Exact restatement of caviar.py:315 including the f in {0,1} edge cases: an element
    whose expression is NaN (0*inf) contributes 0, one that is -inf contributes -DBL_MAX.
    """
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        reg = S * np.log(f) + (cnt - S) * np.log(1 - f)
        one = -DBL_MAX * (cnt - n1)      # f == 1: lam<1 elements are -inf -> -DBL_MAX
        zero = -DBL_MAX * (cnt - n0)     # f == 0: lam>0 elements are -inf -> -DBL_MAX
        out = np.where(f == 1.0, one, np.where(f == 0.0, zero, reg))
        out = np.where(np.isnan(f), 0.0, out)
        out = np.where(cnt == 0, 0.0, out)
    return out


def _nll_reduced(stats, phi, phi_prior, prec, t):
    cnt, S, n0, n1, pvals = stats
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        f = sigmoid(phi[:, 0:1] * pvals[None, :] - phi[:, 1:2])
        ll = np.sum(_group_loglik(f, cnt, S, n0, n1), axis=1)
        d = phi - phi_prior
        quad = 0.5 * np.einsum("ni,nij,nj->n", d, prec, d)
        return -ll - np.sum(np.log(phi), axis=1) / t + quad


def laplace_approx(lam, I, phi_prior, phi_cov_prior, form, powers=None, t=1e1,
                   backtrack_alpha=0.25, backtrack_beta=0.5, max_backtrack_iters=40, newton_steps=10):
    """caviar.py:253-310, vectorised over neurons with per-neuron loop predicates
    (what vmap of scan/while_loop does).  Returns phi (N,2), phi_cov (N,2,2)."""
    N = lam.shape[0]
    phi_prior = np.asarray(phi_prior, dtype=np.float64)
    prec = np.linalg.inv(np.asarray(phi_cov_prior, dtype=np.float64))
    phi = phi_prior.copy()
    cov = np.zeros((N, 2, 2))
    if form == "reduced":
        stats = suff_stats(lam, I, powers)
        cnt, S, n0, n1, pvals = stats
        nll = lambda ph: _nll_reduced(stats, ph, phi_prior, prec, t)
    else:
        nll = lambda ph: _nll_literal(lam, ph, phi_prior, prec, I, t)
    for _ in range(newton_steps):
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            if form == "reduced":
                f = sigmoid(phi[:, 0:1] * pvals[None, :] - phi[:, 1:2])
                r = S - cnt * f
                wgt = cnt * f * (1 - f)
                j1 = -np.sum(pvals * r, 1)
                j2 = np.sum(r, 1)
                h11 = np.sum(pvals ** 2 * wgt, 1)
                h12 = -np.sum(pvals * wgt, 1)
                h22 = np.sum(wgt, 1)
            else:
                f = sigmoid(phi[:, 0:1] * I - phi[:, 1:2])
                j1 = -np.sum(I * (lam - f), 1)
                j2 = np.sum(lam - f, 1)
                h11 = np.sum(I ** 2 * f * (1 - f), 1)
                h12 = -np.sum(I * f * (1 - f), 1)
                h22 = np.sum(f * (1 - f), 1)
            J = np.stack([j1, j2], 1) + np.einsum("nij,nj->ni", prec, phi - phi_prior) - 1 / (t * phi)
            H = np.zeros((N, 2, 2))
            H[:, 0, 0], H[:, 0, 1], H[:, 1, 0], H[:, 1, 1] = h11, h12, h12, h22
            H = H + prec
            H[:, 0, 0] += 1 / (t * phi[:, 0] ** 2)
            H[:, 1, 1] += 1 / (t * phi[:, 1] ** 2)
            det = H[:, 0, 0] * H[:, 1, 1] - H[:, 0, 1] * H[:, 1, 0]
            Hinv = np.empty_like(H)
            Hinv[:, 0, 0], Hinv[:, 1, 1] = H[:, 1, 1] / det, H[:, 0, 0] / det
            Hinv[:, 0, 1], Hinv[:, 1, 0] = -H[:, 0, 1] / det, -H[:, 1, 0] / det
            v = -np.einsum("nij,nj->ni", Hinv, J)
            step = np.ones(N)
            base = nll(phi)
            Jv = np.sum(J * v, 1)
            lhs = nll(phi + step[:, None] * v)
            rhs = base + backtrack_alpha * step * Jv
            bt = np.zeros(N, dtype=int)
            while True:
                go = (bt < max_backtrack_iters) & (np.isnan(lhs) | (lhs > rhs))
                if not go.any():
                    break
                bt = bt + go
                step = np.where(go, step * backtrack_beta, step)
                lhs_new = nll(phi + step[:, None] * v)
                lhs = np.where(go, lhs_new, lhs)
                rhs = np.where(go, base + backtrack_alpha * step * Jv, rhs)
            cov = Hinv
            phi = phi + step[:, None] * v
    return phi, cov


# --------------------------------------------------------------------------- a8
def estimate_spont_act_soft_thresh(y, mu, lam, it, max_iters, z, pen, mask, scale_factor, tol,
                                   spont_orthogonality):
    """caviar.py:146-163 with the call-site carry of caviar.py:86-87 (loop counter = outer `it`,
    err0 = sum(y))."""
    err = np.sum(y)
    j = it
    resid = y - lam.T @ mu
    blocked = np.any(lam >= spont_orthogonality, axis=0)
    ysq = np.sum(np.square(y)) + 1e-5
    while j < max_iters and err > tol:
        zz = np.where(resid < pen, 0.0, resid - pen)
        zz = np.where(zz < 0.0, 0.0, zz)
        zz = np.where(blocked, 0.0, zz)
        zz = zz * mask
        z = zz
        err = np.sum(np.square(resid - z)) / ysq
        j += 1
        pen *= scale_factor
    return z


# --------------------------------------------------------------------------- a9
def reconnect_spont_cells(y, stim_matrix, lam, mu, beta, z, minimax_spk_prob=0.3, minimum_spike_count=3,
                          log=None):
    """caviar.py:102-144."""
    disc_cells = np.where(mu == 0.0)[0]
    powers = np.unique(stim_matrix)[1:]
    z = np.array(z)
    mu, beta, lam = mu.copy(), beta.copy(), lam.copy()
    while len(disc_cells) > 0:
        if len(np.where(z)[0]) > minimum_spike_count:
            stim_locs = [np.where(z[np.where(stim_matrix[n])[0]])[0] for n in disc_cells]
            focus_indx = int(np.argmax([len(sl) for sl in stim_locs]))
            focus = disc_cells[focus_indx]
            srates = np.zeros_like(powers)
            spike_count = 0
            for i, p in enumerate(powers):
                z_locs = np.where(stim_matrix[focus] == p)[0]
                if len(z_locs) > 0:
                    srates[i] = np.mean(z[z_locs] != 0)
                    spike_count += np.sum(z[z_locs] != 0)
            pava = isotonic_regression(srates)[-1]
            if pava >= minimax_spk_prob and spike_count >= minimum_spike_count:
                z_locs = np.intersect1d(np.where(stim_matrix[focus])[0], np.where(z)[0])
                mu[focus] = np.mean(z[z_locs])
                with np.errstate(invalid="ignore", divide="ignore"):
                    beta[focus] = sem(z[z_locs]) if len(z_locs) > 1 else np.nan
                lam[focus, z_locs] = 1.0
                z[z_locs] = 0.0
                if log is not None:
                    log.append((int(focus), float(pava), int(spike_count)))
            disc_cells = np.delete(disc_cells, focus_indx)
        else:
            break
    return mu, beta, lam, z


# --------------------------------------------------------------------------- driver
def caviar(y_psc, I, mu_prior, beta_prior, shape_prior, rate_prior, phi_prior, phi_cov_prior,
           iters=50, num_mc_samples=100, seed=0, y_xcorr_thresh=1e-2, minimum_spike_count=3,
           delay_spont_est=1, msrmp=0.3, scale_factor=0.75, penalty=5e0, save_histories=False,
           max_backtrack_iters=20, tol=0.05, spont_orthogonality=0.1, fn_scan=True,
           form="reduced", trace=None):
    """caviar.py:20-100.  Returns the reference's 17-tuple.  `form`/`trace` are oracle-only."""
    y, lam_mask, lam, I, powers = prologue(y_psc, I, y_xcorr_thresh)
    spont_rate = 0.0
    N = mu_prior.shape[0]
    K = y.shape[0]
    mu_prior = np.asarray(mu_prior, dtype=np.float64)
    beta_prior = np.asarray(beta_prior, dtype=np.float64)
    mu = mu_prior.copy()
    beta = beta_prior.copy()
    shape = float(shape_prior)
    rate = float(rate_prior)
    phi = np.array(phi_prior, dtype=np.float64)
    phi_cov = np.array(phi_cov_prior, dtype=np.float64)
    z = np.zeros(K)
    receptive_fields = None

    if save_histories:
        hist = [np.zeros((iters, N)), np.zeros((iters, N)), np.zeros((iters, N, K)), np.zeros((iters, K)),
                np.zeros((iters, K)), np.zeros((iters, N, 2)), np.zeros((iters, N, 2, 2)), np.zeros((iters, K))]
    else:
        hist = [None] * 8

    key = prng.prng_key(seed)
    solve = block_update_mu if form == "literal" else block_update_mu_active
    for it in range(iters):
        mu, beta = solve(y, lam, shape, rate, mu_prior, beta_prior)
        lam, key = update_lam(y, I, mu, beta, lam, shape, rate, phi, phi_cov, lam_mask, key, num_mc_samples,
                              powers, minimum_spike_count, msrmp + spont_rate, it, delay_spont_est, form,
                              trace["decisions"] if trace is not None else None)
        shape, rate = update_sigma(y, mu, beta, lam, shape_prior, rate_prior)
        phi, phi_cov = laplace_approx(lam, I, phi_prior, phi_cov_prior, form, powers)
        key = prng.split(key)[1]                     # caviar.py:250-251,304: keys[-1] == split(key)[1]
        z = estimate_spont_act_soft_thresh(y, mu, lam, it, max_backtrack_iters, z, penalty, lam_mask,
                                           scale_factor, tol, spont_orthogonality)
        spont_rate = np.mean(z != 0.0)
        if save_histories:
            for h, pa in zip(hist, [mu, beta, lam, shape, rate, phi, phi_cov, z]):
                h[it] = pa
        if trace is not None:
            trace.setdefault("iters", []).append(dict(mu=mu.copy(), beta=beta.copy(), shape=shape, rate=rate,
                                                      phi=phi.copy(), phi_cov=phi_cov.copy(), z=z.copy(),
                                                      spont_rate=spont_rate, lam_sum=lam.sum(1)))
    if fn_scan:
        log = trace.setdefault("reconnect", []) if trace is not None else None
        mu, beta, lam, z = reconnect_spont_cells(y, I, lam, mu, beta, z, minimax_spk_prob=msrmp,
                                                 minimum_spike_count=minimum_spike_count, log=log)
        phi, phi_cov = laplace_approx(lam, I, phi_prior, phi_cov_prior, form, powers)
    return (mu, beta, lam, np.float64(shape), np.float64(rate), phi, phi_cov, z, receptive_fields, *hist)


def default_priors(N):
    """model.py:24-31."""
    return dict(alpha=0.25 * np.ones(N), phi=np.c_[1e-1 * np.ones(N), 5e0 * np.ones(N)],
                phi_cov=np.array([np.array([[1e-1, 0], [0, 1e0]]) for _ in range(N)]),
                mu=np.zeros(N), beta=1e1 * np.ones(N), shape=1.0, rate=1e-1)


def fit(psc, stim, priors=None, **fit_options):
    """Model(N).fit(psc, stim, method='caviar', fit_options) -> state dict (model.py:104-162)."""
    N = stim.shape[0]
    pr = default_priors(N)
    if priors:
        pr.update(priors)
    res = caviar(psc, stim, pr["mu"], pr["beta"], pr["shape"], pr["rate"], pr["phi"], pr["phi_cov"], **fit_options)
    keys = ["mu", "beta", "lam", "shape", "rate", "phi", "phi_cov", "z"]
    state = dict(pr)
    state.update({k: np.array(v) for k, v in zip(keys, res[:8])})
    return state
