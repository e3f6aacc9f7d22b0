"""Host-side behaviour of the drop-in classes that does not need a GPU."""
import numpy as np
import pytest
import torch

from circuitmap_b200 import Model, NeuralDemixer, optimise
from circuitmap_b200.neural_waveform_demixing import state_dict_keys, load_weights, random_weights
from tests.conftest import GOLDEN
import os

no_gpu = not torch.cuda.is_available()


def test_model_default_priors_match_reference():
    m = Model(5)                                                      # model.py:24-34
    assert np.array_equal(m.state["alpha"], 0.25 * np.ones(5))
    assert np.array_equal(m.state["phi"], np.c_[0.1 * np.ones(5), 5.0 * np.ones(5)])
    assert m.state["phi_cov"].shape == (5, 2, 2) and m.state["phi_cov"][0].tolist() == [[0.1, 0], [0, 1.0]]
    assert np.array_equal(m.state["mu"], np.zeros(5)) and np.array_equal(m.state["beta"], 10 * np.ones(5))
    assert m.state["shape"] == 1.0 and m.state["rate"] == 0.1
    pri = {"beta": 3 * np.ones(5)}
    m2 = Model(5, priors=pri)
    assert "mu" in pri and m2.state["beta"][0] == 3                  # caller's dict is filled by setdefault


def test_unknown_method_raises_bare_exception():
    with pytest.raises(Exception):
        Model(3).fit(np.zeros((4, 900)), np.zeros((3, 4)), method="nope")


def test_unknown_fit_option_is_a_typeerror():
    # the live spelling is `msrmp`; stale scripts pass `minimax_spk_prob` and get a TypeError (caviar.py:21-23)
    with pytest.raises(TypeError):
        optimise._validate_options({"minimax_spk_prob": 0.3})
    assert optimise._validate_options({"iters": 3})["iters"] == 3


@pytest.mark.skipif(not no_gpu, reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_cuda():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        NeuralDemixer(path=os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Model(3).fit(np.ones((4, 900)), np.ones((3, 4)))


def test_weight_loading_and_key_order():
    keys = state_dict_keys()
    assert len(keys) == 54 and keys[0] == "dblock1.conv.weight" and keys[-1] == "conv.bn.running_var"
    w = load_weights(os.path.join(GOLDEN, "nwd_ie_ChroME2f_weights.npz"))
    assert w["ublock4.deconv.weight"].shape == (32, 4, 32) and w["conv.conv.weight"].dtype == np.float32
    r = random_weights()
    assert all(r[k].shape == w[k].shape for k in keys)
    # 76,018 checkpoint elements = these 54 tensors + 9 num_batches_tracked scalars
    assert sum(v.size for v in w.values()) + 9 == 76018


def test_product_never_imports_the_oracle():
    import circuitmap_b200, pathlib
    for f in pathlib.Path(circuitmap_b200.__file__).parent.rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f


def test_lightning_checkpoints_load_without_lightning():
    """`NeuralDemixer(path=...ckpt)` reads the reference's Lightning checkpoints with plain torch.load (no
    pytorch_lightning needed) and yields the tensors of the committed golden fixtures.  Runs where the reference
    checkout is mounted (this container); skipped elsewhere."""
    import glob

    import numpy as np
    import pytest

    from circuitmap_b200.neural_waveform_demixing import load_weights, state_dict_keys
    from tests.conftest import GOLDEN
    ckpts = sorted(glob.glob("/root/reference/demixers/*.ckpt"))
    if not ckpts:
        pytest.skip("reference checkpoints not mounted")
    for f in ckpts:
        w = load_weights(f)
        assert list(w) == state_dict_keys() and all(v.dtype == np.float32 for v in w.values())
    for name in ("nwd_ie_ChroME2f", "nwd_ee_ChroME1"):
        w = load_weights("/root/reference/demixers/%s.ckpt" % name)
        g = dict(np.load(os.path.join(GOLDEN, name + "_weights.npz")))
        assert all(np.array_equal(w[k], g[k]) for k in w)


def test_pack_stim_u8_host_helper():
    """cm_pack_stim_u8 (csrc/stim.cu, host only): powers = np.unique(I)[1:] (caviar.py:42), nnz, uint8 codes."""
    from circuitmap_b200 import optimise
    rng = np.random.default_rng(0)
    for dtype in (np.float64, np.float32):
        I = np.zeros((37, 5000), dtype=dtype)
        pw = np.array([45, 55, 65, 70.5], dtype=dtype)
        sel = rng.random(I.shape) < 0.03
        I[sel] = pw[rng.integers(0, 4, int(sel.sum()))]
        codes, powers, nnz = optimise.pack_stim_host(I)
        assert np.array_equal(powers, np.unique(I)[1:].astype(np.float64)) and nnz == np.count_nonzero(I)
        want = np.zeros(I.shape, np.uint8)
        for i, p in enumerate(powers):
            want[I == dtype(p)] = i + 1
        assert np.array_equal(codes.numpy(), want)
    # a design without zeros: the smallest value is dropped like the reference does (A.3 #8); its entries become code 255
    J = np.full((3, 40), 45.0); J[1, :5] = 65.0
    codes, powers, nnz = optimise.pack_stim_host(J)
    assert powers.tolist() == [65.0] and nnz == 120 and set(np.unique(codes.numpy())) == {1, 255}
    with pytest.raises(RuntimeError, match="distinct stimulus powers"):
        optimise.pack_stim_host(np.arange(1.0, 41.0).reshape(2, 20))
    with pytest.raises(RuntimeError, match="negative"):
        optimise.pack_stim_host(np.array([[0.0, -1.0, 45.0]]))
