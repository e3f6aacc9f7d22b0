"""Pins the NWD oracle against golden vectors produced by the UNMODIFIED reference module (oracle/make_golden.py)."""
import os

import numpy as np

from oracle import nwd
from tests.conftest import GOLDEN


def _load():
    g = np.load(os.path.join(GOLDEN, "nwd_golden.npz"))
    sd = dict(np.load(os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz")))
    return g, sd


def test_forward_np_fp64_matches_reference_double_network():
    g, sd = _load()
    tr = g["traces"]
    out = nwd.forward_np(tr / tr.max(1)[:, None], nwd.fold_bn(sd))
    assert np.max(np.abs(out - g["net_out_f64"])) < 1e-12


def test_torch_restatement_is_bit_identical_to_reference_module():
    g, sd = _load()
    t = nwd.TorchNWD(sd)
    assert np.array_equal(t.demix(g["traces"].copy()), g["out"])
    assert np.array_equal(t.demix(g["traces"].copy(), monotone_start=900), g["out_nofilt"])


def test_demix_np_and_monotone_filter():
    g, sd = _load()
    d = nwd.demix_np(g["traces"].copy(), nwd.fold_bn(sd))
    tmax = g["traces"].max(1)[:, None]
    assert np.max(np.abs(d - g["out"]) / tmax) < 1e-4                 # fp64 vs the reference's fp32 network
    assert np.all(np.diff(d[:, 499:], axis=1) <= 0)
    a = np.array([[3.0, 1.0, 2.0, 0.5, 4.0]])
    assert nwd.monotone_decay_filter(a.copy(), 2).tolist() == [[3.0, 1.0, 1.0, 0.5, 0.5]]


def test_intermediate_activations_shapes():
    g, sd = _load()
    assert g["act_dblock1"].shape == (2, 16, 387) and g["act_ublock3"].shape == (2, 32, 387)
    assert g["act_ublock4"].shape == (2, 4, 900) and g["act_conv"].shape == (2, 1, 900)
