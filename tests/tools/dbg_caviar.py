"""Debug helper (not a test): per-iteration comparison of the CUDA fit against the oracle trace."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import caviar as oc, simulate as osim
from circuitmap_b200 import Model

N, K, H, seed, iters = [int(x) for x in (sys.argv[1:6] + ["32", "300", "4", "3", "30"][len(sys.argv) - 1:])]
np.random.seed(seed)
sim = osim.simulate(N=N, trials=K, H=H, connection_prob=0.15 if N < 100 else 0.1)
psc, stim = sim["psc"], sim["stim_matrix"]
pr = oc.default_priors(N)
tr = {"decisions": []}
t = time.time()
ref = oc.caviar(psc, stim, pr["mu"], pr["beta"], pr["shape"], pr["rate"], pr["phi"], pr["phi_cov"], iters=iters, seed=1,
                msrmp=0.4, form="reduced", trace=tr)
print("oracle %.2fs" % (time.time() - t))
m = Model(N)
t = time.time()
m.fit(psc, stim, method="caviar", fit_options=dict(iters=iters, seed=1, msrmp=0.4, save_histories=True))
print("gpu fit wall %.3fs" % (time.time() - t))
def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.nanmax(np.abs(a - b) / (np.abs(b) + 1e-9))) if a.size else 0.0
for it in range(iters):
    o = tr["iters"][it]
    h = m.history
    print(it, "mu %.2e beta %.2e shape %.2e rate %.2e phi %.2e cov %.2e z %.2e lamsum %.2e | nconn %d/%d" % (
        rel(h["mu"][it], o["mu"]), rel(h["beta"][it], o["beta"]), rel(h["shape"][it][0], o["shape"]),
        rel(h["rate"][it][0], o["rate"]), rel(h["phi"][it], o["phi"]), rel(h["phi_cov"][it], o["phi_cov"]),
        rel(h["z"][it], o["z"]), rel(h["lam"][it].sum(1), o["lam_sum"]),
        int((h["lam"][it].sum(1) > 0).sum()), int((o["lam_sum"] > 0).sum())))
for i, nm in enumerate(["mu", "beta", "lam", "shape", "rate", "phi", "phi_cov", "z"]):
    print("final", nm, rel(m.state[nm], ref[i]))
print("connected gpu", np.nonzero(m.state["mu"])[0], "oracle", np.nonzero(ref[0])[0], "truth", np.nonzero(sim["weights"])[0])
