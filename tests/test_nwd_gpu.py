"""GPU parity: cm_nwd_forward (through the C ABI / NeuralDemixer) vs the oracle and the reference golden vectors."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu

# fp32 CUDA-core path: differs from the reference's fp32 oneDNN path only by summation order and BN folding.
# torch fp32 vs fp64 of the same network is 2.6e-5 max-abs on unit-normalised traces.
TOL_UNIT = 2e-4   # max-abs on unit-normalised traces


@pytest.fixture(scope="module")
def demixer():
    from circuitmap_b200 import NeuralDemixer
    return NeuralDemixer(path=os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz"))


def test_golden_reference_vectors(demixer):
    g = np.load(os.path.join(GOLDEN, "nwd_golden.npz"))
    traces = g["traces"]
    tmax = traces.max(1)[:, None]
    out = demixer(traces.copy(), verbose=False)
    assert out.dtype == np.float64 and out.shape == traces.shape
    assert np.max(np.abs(out - g["out"]) / tmax) < TOL_UNIT
    out_nf = demixer(traces.copy(), monotone_filter_start=900, verbose=False)
    assert np.max(np.abs(out_nf - g["out_nofilt"]) / tmax) < TOL_UNIT
    # against the fp64 network: bounded by fp32 rounding of the whole net
    assert np.max(np.abs(out_nf / tmax - g["net_out_f64"])) < TOL_UNIT


def test_vs_oracle_seeded(demixer):
    from oracle import nwd as onwd
    from oracle.make_golden import synth_traces
    sd = dict(np.load(os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz")))
    folded = onwd.fold_bn(sd)
    traces = synth_traces(300, seed=11)
    ref = onwd.demix_np(traces.copy(), folded)
    out = demixer(traces.copy(), verbose=False)
    tmax = traces.max(1)[:, None]
    assert np.max(np.abs(out - ref) / tmax) < TOL_UNIT
    # monotone property beyond sample 500 and idempotence of the filter
    assert np.all(np.diff(out[:, 499:], axis=1) <= 0)


def test_fp32_io_and_stats(demixer):
    import torch
    from oracle.make_golden import synth_traces
    traces = synth_traces(64, seed=5)
    x64 = torch.from_numpy(traces).cuda()
    o64, y, ss = demixer.forward_device(x64, stats=True)
    o32 = demixer.forward_device(x64.float())
    tmax = torch.from_numpy(traces.max(1)[:, None]).cuda()
    assert torch.max(torch.abs(o32.double() - o64) / tmax).item() < TOL_UNIT
    ref_y = o64.sum(1) - 0.5 * (o64[:, 0] + o64[:, -1])
    assert torch.allclose(y, ref_y, rtol=1e-12, atol=1e-12)
    assert torch.allclose(ss, (o64 * o64).sum(1), rtol=1e-12, atol=1e-14)


def test_ragged_and_edge_batches(demixer):
    from oracle.make_golden import synth_traces
    traces = synth_traces(3, seed=2)
    full = demixer(traces.copy(), verbose=False)
    one = demixer(traces[:1].copy(), verbose=False)
    assert one.shape == (1, 900)
    assert np.array_equal(one[0], full[0])              # batch-size independence, bit exact
    empty = demixer(np.zeros((0, 900)), verbose=False)
    assert empty.shape == (0, 900)
    with pytest.raises(RuntimeError):
        demixer(np.ones((2, 800)), verbose=False)       # T != 900 is rejected loudly


def test_tensor_core_path_within_stated_bound():
    """precision='tf32': tcgen05 implicit-GEMM convolutions.  Bound stated in include/circuitmap_b200.h:
    max-abs <= 2e-2 and relative L2 <= 3e-3 on unit-normalised traces, against the fp64 oracle."""
    from circuitmap_b200 import NeuralDemixer
    from oracle import nwd as onwd
    from oracle.make_golden import synth_traces
    sd = dict(np.load(os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz")))
    folded = onwd.fold_bn(sd)
    traces = synth_traces(400, seed=21)
    tmax = traces.max(1)[:, None]
    ref = onwd.demix_np(traces.copy(), folded, monotone_start=900) / tmax
    dem = NeuralDemixer(path=os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz"), precision="tf32")
    out = dem(traces.copy(), monotone_filter_start=900, verbose=False) / tmax
    err = np.abs(out - ref)
    rel_l2 = np.sqrt((err ** 2).sum(1)) / (np.sqrt((ref ** 2).sum(1)) + 1e-3)
    print("tf32 path: max-abs %.3e, median rel-L2 %.3e, max rel-L2 %.3e" % (err.max(), np.median(rel_l2), rel_l2.max()))
    assert err.max() < 2e-2
    assert np.median(rel_l2) < 3e-3
    assert np.sqrt((err ** 2).sum()) / np.sqrt((ref ** 2).sum()) < 3e-3          # pooled relative L2
    # same call in fp32 mode on the same handle is tighter by two orders of magnitude
    dem.set_precision("fp32")
    out32 = dem(traces.copy(), monotone_filter_start=900, verbose=False) / tmax
    assert np.abs(out32 - ref).max() < TOL_UNIT


# ---- multi-trace fp16 tensor-core path (csrc/nwd_mt.cu) ---------------------------------------------------------------
@pytest.fixture(scope="module")
def demixer16():
    from circuitmap_b200 import NeuralDemixer
    return NeuralDemixer(path=os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz"), precision="fp16")


def test_fp16_multitrace_path_within_stated_bound(demixer16):
    """precision='fp16': every convolution as a widened tcgen05 implicit GEMM, 4 traces per M tile.  Bound stated in
    include/circuitmap_b200.h: max-abs <= 2e-2 and relative L2 <= 3e-3 on unit-normalised traces vs the fp64 oracle;
    also checked against the reference's own golden vectors."""
    from oracle import nwd as onwd
    from oracle.make_golden import synth_traces
    sd = dict(np.load(os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz")))
    folded = onwd.fold_bn(sd)
    traces = synth_traces(403, seed=21)                    # not a multiple of 4: the last pass is ragged
    tmax = traces.max(1)[:, None]
    ref = onwd.demix_np(traces.copy(), folded, monotone_start=900) / tmax
    out = demixer16(traces.copy(), monotone_filter_start=900, verbose=False) / tmax
    err = np.abs(out - ref)
    rel_l2 = np.sqrt((err ** 2).sum(1)) / (np.sqrt((ref ** 2).sum(1)) + 1e-3)
    print("fp16 path: max-abs %.3e, median rel-L2 %.3e, pooled rel-L2 %.3e" % (
        err.max(), np.median(rel_l2), np.sqrt((err ** 2).sum()) / np.sqrt((ref ** 2).sum())))
    assert err.max() < 2e-2
    assert np.median(rel_l2) < 3e-3
    assert np.sqrt((err ** 2).sum()) / np.sqrt((ref ** 2).sum()) < 3e-3
    g = np.load(os.path.join(GOLDEN, "nwd_golden.npz"))
    gt = g["traces"]
    gout = demixer16(gt.copy(), verbose=False)
    assert np.max(np.abs(gout - g["out"]) / gt.max(1)[:, None]) < 2e-2
    assert np.all(np.diff(gout[:, 499:], axis=1) <= 0)      # monotone decay beyond sample 500


def test_fp16_path_rows_are_independent_of_their_tile_mates(demixer16):
    """Several traces share one 128-row MMA tile; a trace's result must not depend on which traces sit next to it,
    on its slot in the pass, or on the batch size (bit exact)."""
    from oracle.make_golden import synth_traces
    traces = synth_traces(37, seed=3)
    full = demixer16(traces.copy(), verbose=False)
    for idx in ([0], [5, 6], [36, 1, 17], list(range(36, -1, -1)), [9] * 5):
        sub = demixer16(traces[idx].copy(), verbose=False)
        assert np.array_equal(sub, full[idx])
    again = demixer16(traces.copy(), verbose=False)
    assert np.array_equal(again, full)                      # run-to-run determinism
    assert demixer16(np.zeros((0, 900)), verbose=False).shape == (0, 900)


def test_fp16_path_io_dtypes_stats_and_large_batch(demixer16):
    import torch
    from oracle.make_golden import synth_traces
    traces = synth_traces(70, seed=5)
    x64 = torch.from_numpy(traces).cuda()
    o64, y, ss = demixer16.forward_device(x64, stats=True)
    o32 = demixer16.forward_device(x64.float())
    tmax = torch.from_numpy(traces.max(1)[:, None]).cuda()
    assert torch.max(torch.abs(o32.double() - o64) / tmax).item() < 1e-3
    ref_y = o64.sum(1) - 0.5 * (o64[:, 0] + o64[:, -1])
    assert torch.allclose(y, ref_y, rtol=1e-12, atol=1e-12)
    assert torch.allclose(ss, (o64 * o64).sum(1), rtol=1e-12, atol=1e-14)
    # full C2 size (20000 traces, > 33 passes per SM): every block of 70 traces repeats the small batch bit for bit
    big = x64.float().repeat(286, 1)[:20000].contiguous()
    ob = demixer16.forward_device(big)
    assert torch.equal(ob[:70], o32) and torch.equal(ob[19950:20000], ob[19950 - 70 * 100:20000 - 70 * 100])
    assert torch.isfinite(ob).all()


def test_fp16_path_flags_traces_outside_the_fp16_range(demixer16):
    """fp16 operands: a trace that cannot be normalised into the fp16 range (max 0, non-finite samples, |x / max| > 6e4)
    comes back as an all-NaN row -- loudly, never saturated -- and does not disturb the other traces of its tile."""
    from oracle.make_golden import synth_traces
    traces = synth_traces(8, seed=9)
    good = demixer16(traces.copy(), verbose=False)
    bad = traces.copy()
    bad[1] = 0.0                      # max == 0 (the reference divides by zero here, too)
    bad[2, 100] = np.nan
    bad[5] = -np.abs(bad[5]) * 1e6    # max is tiny and negative, min is huge
    bad[5, 0] = -1e-3
    out = demixer16(bad, verbose=False)
    for i in (1, 2, 5):
        assert np.isnan(out[i]).all()
    for i in (0, 3, 4, 6, 7):
        assert np.array_equal(out[i], good[i])
    # an activation that overflows fp16 inside the network is flagged the same way (never saturated silently)
    from circuitmap_b200 import NeuralDemixer
    from circuitmap_b200.neural_waveform_demixing import random_weights
    sd = random_weights(seed=1)
    sd["dblock1.bn.weight"] = sd["dblock1.bn.weight"] * 3e5      # weights still fit fp16 (<= 5.3e4), activations do not
    big = NeuralDemixer(state_dict=sd, precision="fp16")
    assert np.isnan(big(traces.copy(), verbose=False)).all()
    big.set_precision("fp32")
    assert np.isfinite(big(traces.copy(), verbose=False)).all()
    sd["dblock1.bn.weight"] = sd["dblock1.bn.weight"] * 10       # now the folded weights themselves overflow fp16
    with pytest.raises(RuntimeError):
        NeuralDemixer(state_dict=sd, precision="fp16")
    assert np.isfinite(NeuralDemixer(state_dict=sd, precision="fp32")(traces.copy(), verbose=False)).all()


@pytest.mark.parametrize("weights", ["nwd_ee_ChroME1_weights.npz", "random"])
def test_fp16_path_bound_holds_for_other_networks(weights):
    """The stated bound of the fp16 tensor-core path is not specific to one checkpoint: second reference checkpoint
    (excitatory ChroME1 demixer) and a random-init network of the same architecture, against the fp64 oracle."""
    from circuitmap_b200 import NeuralDemixer
    from circuitmap_b200.neural_waveform_demixing import random_weights
    from oracle import nwd as onwd
    from oracle.make_golden import synth_traces
    if weights == "random":
        sd = random_weights(seed=3)
        dem = NeuralDemixer(path=None, precision="fp16")
        dem2 = NeuralDemixer(path=None, precision="fp32")
        assert all(np.array_equal(dem.weights[k], random_weights(seed=0)[k]) for k in dem.weights)   # path=None -> seed 0
        sd = random_weights(seed=0)
    else:
        sd = dict(np.load(os.path.join(GOLDEN, weights)))
        dem = NeuralDemixer(path=os.path.join(GOLDEN, weights), precision="fp16")
        dem2 = NeuralDemixer(path=os.path.join(GOLDEN, weights), precision="fp32")
    folded = onwd.fold_bn(sd)
    traces = synth_traces(200, seed=33)
    tmax = traces.max(1)[:, None]
    ref = onwd.demix_np(traces.copy(), folded, monotone_start=900) / tmax
    out = dem(traces.copy(), monotone_filter_start=900, verbose=False) / tmax
    out32 = dem2(traces.copy(), monotone_filter_start=900, verbose=False) / tmax
    scale = max(1.0, np.abs(ref).max())                      # a random network is not normalised to unit output
    err = np.abs(out - ref).max() / scale
    print("%s: fp16 max-abs %.3e (output scale %.2f), fp32 %.3e" % (weights, err, scale, np.abs(out32 - ref).max() / scale))
    assert err < 2e-2
    assert np.sqrt(((out - ref) ** 2).sum()) / (np.sqrt((ref ** 2).sum()) + 1e-12) < 3e-3
    assert np.abs(out32 - ref).max() / scale < TOL_UNIT
