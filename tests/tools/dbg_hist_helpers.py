import os, sys, numpy as np, torch, io, contextlib
sys.path.insert(0, ".")
from oracle import simulate as osim
from circuitmap_b200 import Model
sim = osim.simulate_fast(N=300, K=3000, H=10, seed=3)
res = {}
for h in ("0", "15"):
    os.environ["CM_CAVIAR_HELPERS"] = h
    m = Model(300)
    with contextlib.redirect_stdout(io.StringIO()):
        m.fit(sim["psc"], sim["stim_matrix"], method="caviar", fit_options=dict(iters=8, msrmp=0.4, seed=2, save_histories=True))
    res[h] = m
a, b = res["0"], res["15"]
ok = all(np.array_equal(a.state[k], b.state[k], equal_nan=True) for k in ("mu", "beta", "lam", "phi", "phi_cov", "z"))
okh = all(np.array_equal(a.history[k], b.history[k], equal_nan=True) for k in a.history)
print("state equal:", ok, "history equal:", okh, "hist shapes", {k: v.shape for k, v in a.history.items()})
