"""Generate the golden fixtures under tests/golden/ (run in the BUILD container only).

    python -m oracle.make_golden [/root/reference]

NWD: imports the UNMODIFIED reference module circuitmap/neural_waveform_demixing.py from the
read-only reference checkout through a minimal `pytorch_lightning` shim (the package is not
installed; the shim only supplies LightningModule := nn.Module + load_from_checkpoint), runs
NeuralDemixer / NWDUNet on seeded synthetic traces and stores inputs, outputs and intermediate
activations.  The checkpoint tensors of demixers/nwd_ie_ChroME2f.ckpt are stored as a plain .npz
(weights are data fixtures, not source) so that tests and bench.py can run on the GPU box where
/root/reference does not exist.

CAVIaR: JAX is absent, so the UNMODIFIED reference package (circuitmap/model.py, optimise/caviar.py,
optimise/pava.py, simulation.py) is imported through oracle/jax_shim.py -- a NumPy-backed stand-in
for the few JAX symbols those files import -- and executed here:
  tests/golden/caviar_ref_tiny_N32_K300.npz   simulate(seed 3) -> Model.fit(iters 30), full histories
  tests/golden/caviar_ref_C1_seed0.npz        C1 (N=100, K=2000, H=10, seed 0) -> Model.fit(iters 50,
                                              seed 1, msrmp 0.4) as scripts/run_simulations.py:54-61
  tests/golden/caviar_ref_reconnect.npz       reconnect_spont_cells (caviar.py:102-144) on a constructed
                                              state: 4 sequential reconnects, a first-arg-max tie, a
                                              single-sample sem() NaN, a rejected cell
These are outputs of the reference's own source text; the PRNG underneath is still the restated
threefry of oracle/prng.py (pinned to Random123 + JAX's own known answers, not to a live JAX).
The oracle's own outputs on the tiny map are also kept as a regression pin of the restatement.

    python -m oracle.make_golden [/root/reference] [nwd|caviar|all]
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_reference_nwd(ref_root):
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        @classmethod
        def _load(cls, path):
            m = cls()
            ck = torch.load(path, map_location="cpu", weights_only=True)
            m.load_state_dict(ck["state_dict"])
            return m

        def load_from_checkpoint(self, path):       # reference calls it on an instance (nwd.py:27)
            return type(self)._load(path)

        def log(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    pl.Trainer = object
    sys.modules["pytorch_lightning"] = pl
    spec = importlib.util.spec_from_file_location(
        "ref_nwd", os.path.join(ref_root, "circuitmap", "neural_waveform_demixing.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synth_traces(K, seed, T=900):
    """PSC-like traces: bi-exponential events + smooth + iid noise (cf. simulation.py:25-29 scales)."""
    rng = np.random.default_rng(seed)
    t = np.arange(T)[None, :]
    tr = rng.uniform(25, 60, (K, 1))
    td = tr + rng.uniform(75, 250, (K, 1))
    d = rng.uniform(100, 400, (K, 1))
    amp = rng.uniform(0.0, 40.0, (K, 1)) * (rng.random((K, 1)) < 0.7)
    with np.errstate(over="ignore"):
        ev = (np.exp(-(t - d) / td) - np.exp(-(t - d) / tr)) * (t > d)
    ev = ev / (ev.sum(1, keepdims=True) + 1e-5) * amp
    d2 = rng.uniform(-300, 850, (K, 1))
    with np.errstate(over="ignore"):
        ev2 = (np.exp(-(t - d2) / td) - np.exp(-(t - d2) / tr)) * (t > d2)
    ev2 = ev2 / (np.abs(ev2).sum(1, keepdims=True) + 1e-5) * rng.uniform(0, 20, (K, 1)) * (rng.random((K, 1)) < 0.3)
    smooth = np.cumsum(rng.normal(0, 2e-4, (K, T)), axis=1)
    return ev + ev2 + smooth + rng.normal(0, 6e-4, (K, T))


def _quiet(fn, *a, **k):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def reconnect_case():
    """Constructed post-loop state for reconnect_spont_cells (caviar.py:102-144).

    8 neurons, 3 powers, 60 trials.  Neuron 0 stays connected (mu != 0).  Disconnected cells:
      1 and 2  tie on the number of spont events in their trials (first arg-max wins, :117) -> 1 first;
               both pass PAVA -> reconnected in sequence; 2 loses the events it shared with 1;
      5        4 events -> examined second (after 1 took 2's shared events), reconnects;
      3        exactly ONE event among its trials but rate 1/2 at the top power -> reconnects with
               beta = sem(single sample) = NaN (:135) when minimum_spike_count = 1, rejected when it is 3;
      4        events only at the LOWEST power: PAVA pools the rates to 2/9 < msrmp -> rejected;
      6, 7     never see an event.
    """
    N, K = 8, 60
    powers = np.array([45.0, 55.0, 65.0])
    stim = np.zeros((N, K))
    z = np.zeros(K)
    stim[0, 0:6] = [45, 55, 65, 45, 55, 65]
    # cell 1: trials 6..14 (3 per power); events on 5 of them
    stim[1, 6:15] = np.repeat(powers, 3)
    # cell 2: trials 12..20 overlap cell 1's top-power trials 12..14
    stim[2, 12:21] = np.r_[np.repeat(65.0, 3), np.repeat(45.0, 3), np.repeat(55.0, 3)]
    z[[7, 9, 10, 12, 13]] = [3.0, 2.5, 4.0, 6.0, 5.5]          # cell 1: 5 events (1@45, 2@55, 2@65)
    z[[15, 18, 19]] = [1.5, 2.0, 2.2]                           # cell 2: +3 own -> 5 with the shared 12, 13
    stim[3, [22, 23]] = 65.0
    stim[3, [24, 25, 26, 27]] = [45.0, 45.0, 55.0, 55.0]
    z[22] = 7.25                                                # single event, rate 0.5 at the top power
    stim[4, 28:37] = np.repeat(powers, 3)
    z[[28, 29]] = [1.0, 1.2]                                    # 2 of 3 at 45 only -> PAVA pools (2/3, 0, 0) -> 2/9 < msrmp
    z[[58, 59]] = [0.7, 0.9]                                    # events on trials nobody stimulates keep the loop alive (:111)
    stim[5, 37:46] = np.repeat(powers, 3)
    z[[41, 43, 44, 45]] = [2.0, 3.0, 3.5, 4.5]
    stim[6, 46:52] = np.repeat(powers, 2)
    stim[7, 52:58] = np.repeat(powers, 2)
    y = z + 0.01
    mu = np.zeros(N)
    mu[0] = 12.0
    beta = np.full(N, 100.0)
    beta[0] = 0.3
    lam = np.zeros((N, K))
    lam[0, [2, 5]] = [0.9, 0.8]
    return dict(y=y, stim=stim, lam=lam, mu=mu, beta=beta, z=z)


def make_caviar_reference(ref_root, parts=("tiny", "c1", "reconnect")):
    """Run the unmodified reference CAVIaR path through the NumPy-backed jax shim."""
    import hashlib
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import jax_shim
    cm = jax_shim.import_reference(ref_root)
    ref_caviar = sys.modules["circuitmap.optimise.caviar"]
    jnp = sys.modules["jax.numpy"]
    hkeys = ["mu", "beta", "lam", "shape", "rate", "phi", "phi_cov", "z"]

    def run(seed_np, simkw, fit_options):
        np.random.seed(seed_np)
        sim = _quiet(cm.simulate, **simkw)
        N = sim["stim_matrix"].shape[0]
        m = cm.Model(N)
        _quiet(m.fit, sim["psc"], sim["stim_matrix"], method="caviar", fit_options=dict(fit_options, save_histories=True))
        return sim, m

    if "tiny" in parts:
        _ref_tiny(run, hkeys)
    if "c1" in parts:
        _ref_c1(run, hkeys, hashlib)
    if "reconnect" in parts:
        _ref_reconnect(ref_caviar, jnp)


def _ref_tiny(run, hkeys):
    # (1) tiny map, everything stored
    fo = dict(iters=30, seed=1, msrmp=0.4)
    sim, m = run(3, dict(N=32, trials=300, H=4, connection_prob=0.15), fo)
    out = dict(psc=np.asarray(sim["psc"]), stim=np.asarray(sim["stim_matrix"]).astype(np.uint8),
               weights=sim["weights"], fit_iters=30, fit_seed=1, fit_msrmp=0.4, sim_seed=3)
    out.update({k: np.asarray(m.state[k], dtype=np.float64) for k in hkeys})
    out.update({"hist_" + k: np.asarray(m.history[k], dtype=np.float64) for k in hkeys})
    np.savez_compressed(os.path.join(GOLD, "caviar_ref_tiny_N32_K300.npz"), **out)
    print("reference tiny: connected", np.nonzero(m.state["mu"])[0], "true", np.nonzero(sim["weights"])[0])



def _ref_c1(run, hkeys, hashlib):
    # (2) C1 as scripts/run_simulations.py:54-61 runs it (raw PSCs); psc is regenerated by the tests from the seed
    fo = dict(iters=50, seed=1, msrmp=0.4)
    sim, m = run(0, dict(N=100, trials=2000, H=10, connection_prob=0.1), fo)
    psc = np.ascontiguousarray(sim["psc"], dtype=np.float64)
    lam_hist = np.asarray(m.history["lam"], dtype=np.float64)
    stim = np.asarray(sim["stim_matrix"])
    out = dict(psc_sha256=np.frombuffer(hashlib.sha256(psc.tobytes()).digest(), dtype=np.uint8),
               psc_y=np.trapezoid(psc, axis=-1), psc_ss=np.sum(psc * psc, axis=-1),
               stim=stim.astype(np.uint8), weights=sim["weights"], fit_iters=50, fit_seed=1, fit_msrmp=0.4, sim_seed=0,
               lam_on_support=np.asarray(m.state["lam"])[stim > 0],
               hist_lam_rowsum=lam_hist.sum(2), hist_lam_rowany=(lam_hist != 0).any(2))
    out.update({k: np.asarray(m.state[k], dtype=np.float64) for k in hkeys if k != "lam"})
    out.update({"hist_" + k: np.asarray(m.history[k], dtype=np.float64) for k in ["mu", "beta", "phi", "phi_cov"]})
    out["hist_rate"] = np.asarray(m.history["rate"], dtype=np.float64)[:, 0]
    out["hist_z_nnz"] = (np.asarray(m.history["z"]) != 0).sum(1)
    out["hist_z_sum"] = np.asarray(m.history["z"]).sum(1)
    np.savez_compressed(os.path.join(GOLD, "caviar_ref_C1_seed0.npz"), **out)
    print("reference C1: connected", np.nonzero(m.state["mu"])[0], "true", np.nonzero(sim["weights"])[0])



def _ref_reconnect(ref_caviar, jnp):
    # (3) reconnect_spont_cells on a constructed state
    rc = reconnect_case()
    res = {}
    for msc in (1, 3):
        mu, beta, lam, z = _quiet(ref_caviar.reconnect_spont_cells, rc["y"], rc["stim"], jnp.array(rc["lam"]),
                                  jnp.array(rc["mu"]), jnp.array(rc["beta"]), jnp.array(rc["z"]),
                                  minimax_spk_prob=0.3, minimum_spike_count=msc)
        res.update({f"msc{msc}_mu": np.asarray(mu), f"msc{msc}_beta": np.asarray(beta),
                    f"msc{msc}_lam": np.asarray(lam), f"msc{msc}_z": np.asarray(z)})
        print(f"reference reconnect (minimum_spike_count={msc}): mu", np.asarray(mu), "beta", np.asarray(beta))
    np.savez_compressed(os.path.join(GOLD, "caviar_ref_reconnect.npz"), **rc, **res)


def main(ref_root="/root/reference", what="all"):
    os.makedirs(GOLD, exist_ok=True)
    if what in ("caviar", "all"):
        make_caviar_reference(ref_root)
    if what in ("tiny", "c1", "reconnect"):
        make_caviar_reference(ref_root, parts=(what,))
    if what != "all" and what != "nwd":
        return
    ref = import_reference_nwd(ref_root)
    torch.manual_seed(0)
    for name in ["nwd_ie_ChroME2f", "nwd_ee_ChroME1"]:
        path = os.path.join(ref_root, "demixers", name + ".ckpt")
        ck = torch.load(path, map_location="cpu", weights_only=True)
        sd = {k: v.numpy() for k, v in ck["state_dict"].items()}
        np.savez_compressed(os.path.join(GOLD, name + "_weights.npz"), **sd)

    path = os.path.join(ref_root, "demixers", "nwd_ie_ChroME2f.ckpt")
    dem = ref.NeuralDemixer(path=path, device="cpu")
    assert not dem.demixer.training
    traces = synth_traces(24, seed=7)
    out = dem(traces.copy(), verbose=False)                               # float64, filtered
    out_nofilt = dem(traces.copy(), monotone_filter_start=900, verbose=False)
    # intermediate activations of the raw network on the first 2 normalised traces
    acts = {}
    net = dem.demixer
    hooks = []
    for nm in ["dblock1", "dblock2", "dblock3", "dblock4", "ublock1", "ublock2", "ublock3", "ublock4", "conv"]:
        hooks.append(getattr(net, nm).register_forward_hook(
            lambda m, i, o, nm=nm: acts.__setitem__(nm, o.detach().numpy().copy())))
    tmax = np.max(traces, axis=1)[:, None]
    x = torch.Tensor((traces / tmax).copy()[:2, None, :])
    net(x)
    for h in hooks:
        h.remove()
    net64 = ref.NWDUNet()
    net64.load_state_dict(net.state_dict())
    net64 = net64.double().eval()
    with torch.no_grad():
        out64 = net64(torch.from_numpy((traces / tmax)[:, None, :])).numpy().squeeze(1)
    np.savez_compressed(os.path.join(GOLD, "nwd_golden.npz"), traces=traces, out=out, out_nofilt=out_nofilt,
                        net_out_f64=out64, **{"act_" + k: v for k, v in acts.items()})
    print("NWD golden written:", traces.shape, out.shape, {k: v.shape for k, v in acts.items()})

    # CAVIaR regression pin (oracle output, NOT a reference golden)
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import caviar as oc, simulate as osim
    np.random.seed(3)
    sim = osim.simulate(N=32, trials=300, H=4, connection_prob=0.15)
    st = oc.fit(sim["psc"], sim["stim_matrix"], iters=30, seed=1, msrmp=0.4)
    np.savez_compressed(os.path.join(GOLD, "caviar_oracle_pin_N32_K300.npz"),
                        psc_y=oc.trapz_rows(sim["psc"]), psc_ss=np.sum(sim["psc"] ** 2, 1),
                        stim=sim["stim_matrix"].astype(np.float32), weights=sim["weights"],
                        **{k: st[k] for k in ["mu", "beta", "shape", "rate", "phi", "phi_cov", "z"]},
                        lam_rowsum=st["lam"].sum(1))
    print("CAVIaR pin written; connected:", np.nonzero(st["mu"])[0], "true:", np.nonzero(sim["weights"])[0])


if __name__ == "__main__":
    main(*sys.argv[1:])
