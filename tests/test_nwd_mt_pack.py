"""Host-side check of the multi-trace tensor-core demixer's GEMM formulation (csrc/nwd_mt.cuh), no GPU needed.

`cm_nwd_mt_pack` returns the fp16 tap tables the kernel streams.  `emulate()` evaluates formula (1) of nwd_mt.cuh,
    D[q][(n, co)] = sum_v sum_ci X[PH q + v][ci] * WS[v + n][ci][co],   output position t = PH q + (PH - 1 - n),
with NumPy on those tables, with the kernel's re-groupings (parity / sample-group split of the first layer, (parity, co)
channels of the stride-2 transposed convolution, sample pairs of the final dilated convolution), and must reproduce the
oracle network (oracle/nwd.py::forward_np, pinned to the reference) up to the fp16 rounding of weights/activations.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.conftest import GOLDEN, ROOT


def _cfg():
    src = open(os.path.join(ROOT, "circuitmap_b200", "csrc", "nwd_mt.cuh")).read()
    rows = re.findall(r"^\s*\{(\d+), (\d+), (\d+), (\d+), (\d+), (\d+), (\d+)\},", src, flags=re.M)
    assert len(rows) == 9
    return [tuple(int(x) for x in r) for r in rows]


def _blob(sd):
    from circuitmap_b200 import _lib
    from circuitmap_b200.neural_waveform_demixing import state_dict_keys
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = C.CDLL(_lib.LIB_PATH)
    lib.cm_nwd_mt_pack.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    arrs = [np.ascontiguousarray(sd[k], dtype=np.float32) for k in state_dict_keys()]
    ptrs = (C.c_void_p * 54)(*[a.ctypes.data_as(C.c_void_p) for a in arrs])
    need = C.c_size_t()
    assert lib.cm_nwd_mt_pack(ptrs, 54, None, 0, C.byref(need)) == 0
    buf = np.zeros(need.value, np.uint8)
    assert lib.cm_nwd_mt_pack(ptrs, 54, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(need)) == 0
    return buf


def _tables(blob, cfg):
    ws, off = [], 0
    for (PH, CIN, COUT, TAPS, Q, SEQ, U) in cfg:
        n = (CIN // 8) * U * COUT * 8
        w = blob[off:off + 2 * n].view(np.float16).astype(np.float64).reshape(CIN // 8, U, COUT, 8)
        ws.append(w.transpose(1, 0, 3, 2).reshape(U, CIN, COUT))          # WS[u][ci][co]
        off += 2 * n
    bias = blob[off:off + 9 * 32 * 4].view(np.float32).astype(np.float64).reshape(9, 32)
    return ws, bias


def _gemm(X, WS, c):
    """formula (1): X (L, CIN) -> out (PH*Q, COUT) indexed by output position t."""
    PH, CIN, COUT, TAPS, Q, SEQ, U = c
    V = TAPS + PH - 1
    Xp = np.zeros((PH * Q + V + PH, CIN))
    Xp[:min(len(X), len(Xp))] = X[:len(Xp)]
    win = np.stack([Xp[PH * q:PH * q + V] for q in range(Q)])            # (Q, V, CIN)
    out = np.zeros((PH * Q, COUT))
    for n in range(PH):
        out[PH * np.arange(Q) + PH - 1 - n] = np.einsum("qvc,vco->qo", win, WS[n:n + V])
    return out


def _pool(x):       # (L, C) AvgPool1d(3, 2)
    Lo = (x.shape[0] - 3) // 2 + 1
    i = 2 * np.arange(Lo)
    return (x[i] + x[i + 1] + x[i + 2]) / 3.0


def _interp(x, size):
    from oracle.nwd import interp_linear
    return interp_linear(x.T, size).T


def emulate(x, ws, bias, cfg, rnd=lambda a: a, keep=None):
    """x (900,) unit-normalised -> (900,) network output, following csrc/nwd_mt.cu step by step.
    `keep` (dict) receives the intermediate activations."""
    relu = lambda a: np.maximum(a, 0)
    p1 = rnd(_pool(x[:, None])[:, 0])                                     # 449
    enc1 = np.zeros((387, 16))
    for p in (0, 1):                                                      # d1: parity sequences, groups of 8 samples
        seq = np.zeros(8 * (29 + 7))
        seq[:len(p1[p::2])] = p1[p::2]
        D = _gemm(seq.reshape(-1, 8), ws[0], cfg[0])                      # (29, 128): [sg][16 m + co]
        for m in range(8):
            s = 8 * np.arange(29) + m
            t = 2 * s + p
            ok = t < 387
            enc1[t[ok]] = relu(D[ok, 16 * m:16 * m + 16] + bias[0, :16])
    enc1 = rnd(enc1)
    enc2 = rnd(relu(_gemm(rnd(_pool(enc1)), ws[1], cfg[1])[:162] + bias[1, :16]))
    enc3 = rnd(relu(_gemm(rnd(_pool(enc2)), ws[2], cfg[2])[:65] + bias[2, :32]))
    enc4 = rnd(relu(_gemm(rnd(_pool(enc3)), ws[3], cfg[3])[:17] + bias[3, :32]))
    pad = lambda a, k: np.pad(a, ((k, k), (0, 0)))
    raw1 = rnd(relu(_gemm(pad(enc4, 15), ws[4], cfg[4])[:32] + bias[4, :16]))
    dec1 = np.concatenate([rnd(_interp(raw1, 65)), enc3], 1)
    raw2 = rnd(relu(_gemm(pad(dec1, 15), ws[5], cfg[5])[:80] + bias[5, :16]))
    dec2 = np.concatenate([rnd(_interp(raw2, 162)), enc2], 1)
    raw3 = rnd(relu(_gemm(pad(dec2, 31), ws[6], cfg[6])[:193] + bias[6, :16]))
    dec3 = np.concatenate([rnd(_interp(raw3, 387)), enc1], 1)
    D = _gemm(pad(dec3, 15), ws[7], cfg[7])[:402]                         # [i][4 par + co]
    raw4 = np.zeros((804, 4))
    raw4[0::2] = relu(D[:, 0:4] + bias[7, :4])
    raw4[1::2] = relu(D[:, 4:8] + bias[7, :4])
    raw4 = rnd(raw4)
    h = rnd(_interp(raw4, 900))
    hp = np.pad(h, ((255, 255), (0, 0)))                                  # 1410
    out = np.zeros(900)
    for p in (0, 1):
        xs = hp[p::2]                                                     # 705 x 4
        xs = np.concatenate([xs, np.zeros((2 * 400 - len(xs), 4))])
        pairs = xs.reshape(-1, 8)                                         # [s][4 e + c]
        D = _gemm(pairs, ws[8], cfg[8])                                   # [sigma][co = inner parity]
        for co in (0, 1):
            sig = np.arange(D.shape[0])
            t = 4 * sig + 2 * co + p
            ok = t < 900
            out[t[ok]] = relu(D[ok, co] + bias[8, 0])
    if keep is not None:
        keep.update(p1=p1, enc1=enc1, enc2=enc2, enc3=enc3, enc4=enc4, raw1=raw1, dec1=dec1, raw2=raw2, dec2=dec2, raw3=raw3,
                    dec3=dec3, raw4=raw4, h=h, out=out)
    return out


@pytest.fixture(scope="module", params=["nwd_ie_ChroME2f_weights.npz", "nwd_ee_ChroME1_weights.npz"])
def setup(request):
    sd = dict(np.load(os.path.join(GOLDEN, request.param)))
    cfg = _cfg()
    ws, bias = _tables(_blob(sd), cfg)
    return sd, cfg, ws, bias


def test_plan_is_consistent(setup):
    _, cfg, ws, _ = setup
    lens_in = [None, 193, 80, 32, 47, 95, 224, 417, 353]
    for l, (PH, CIN, COUT, TAPS, Q, SEQ, U) in enumerate(cfg):
        assert U >= TAPS + 2 * PH - 2 and 16 <= PH * COUT <= 256 and (PH * COUT) % 16 == 0
        if lens_in[l]:
            assert PH * Q >= lens_in[l]
        # zero guard taps on both sides of the table
        assert not ws[l][:PH - 1].any() and not ws[l][PH - 1 + TAPS:].any()


def test_formulation_matches_oracle_network(setup):
    from oracle import nwd as onwd
    from oracle.make_golden import synth_traces
    sd, cfg, ws, bias = setup
    folded = onwd.fold_bn(sd)
    traces = synth_traces(3, seed=4)
    x = traces / traces.max(1)[:, None]
    ref = onwd.forward_np(x, folded)
    for i in range(len(x)):
        out = emulate(x[i], ws, bias, cfg)
        # exact formulation, fp16-rounded weights only
        assert np.max(np.abs(out - ref[i])) < 2e-3, np.max(np.abs(out - ref[i]))
    # with every stored activation rounded to fp16 as the kernel does: the stated bound of the tensor-core path
    h = lambda a: a.astype(np.float16).astype(np.float64)
    out = emulate(x[0], ws, bias, cfg, rnd=h)
    assert np.max(np.abs(out - ref[0])) < 2e-2
