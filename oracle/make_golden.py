"""Generate the golden fixtures under tests/golden/ (run in the BUILD container only).

    python -m oracle.make_golden [/root/reference]

NWD: imports the UNMODIFIED reference module circuitmap/neural_waveform_demixing.py from the
read-only reference checkout through a minimal `pytorch_lightning` shim (the package is not
installed; the shim only supplies LightningModule := nn.Module + load_from_checkpoint), runs
NeuralDemixer / NWDUNet on seeded synthetic traces and stores inputs, outputs and intermediate
activations.  The checkpoint tensors of demixers/nwd_ie_ChroME2f.ckpt are stored as a plain .npz
(weights are data fixtures, not source) so that tests and bench.py can run on the GPU box where
/root/reference does not exist.

CAVIaR: JAX is absent, so no reference run is possible; the oracle's own outputs on a tiny seeded
map are stored as a regression pin of the restatement (NOT a reference golden; parity unpinned).
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


def main(ref_root="/root/reference"):
    os.makedirs(GOLD, exist_ok=True)
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
