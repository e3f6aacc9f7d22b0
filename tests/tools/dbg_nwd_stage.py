"""Debug helper (not a test): dump CTA 0's shared memory at every stage of the multi-trace demixer and compare each
activation buffer with the NumPy emulation of tests/test_nwd_mt_pack.py."""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from circuitmap_b200 import NeuralDemixer, _lib
from oracle.make_golden import synth_traces
from tests import test_nwd_mt_pack as E
lib = _lib.load()
lib.cm_nwd_mt_debug_dump.argtypes = [C.c_void_p, C.c_int]
W = "tests/golden/nwd_ie_ChroME2f_weights.npz"
sd = dict(np.load(W)); cfg = E._cfg(); ws, bias = E._tables(E._blob(sd), cfg)
G = 4
traces = synth_traces(G, seed=21)
x = traces / traces.max(1)[:, None]
h16 = lambda a: a.astype(np.float16).astype(np.float64)
keeps = []
for g in range(G):
    k = {}; E.emulate(x[g].astype(np.float32).astype(np.float64), ws, bias, cfg, rnd=h16, keep=k); keeps.append(k)

# shared-memory map (mirror of csrc/nwd_mt.cu)
DEC3_RL, DEC2_RL, DEC1_RL = G * 27 + 3, G * 28 + 5, G * 24 + 10
DEC3_PL, DEC2_PL, DEC1_PL = 16 * DEC3_RL * 16, 8 * DEC2_RL * 16, 4 * DEC1_RL * 16
A_LO = 0; A_END = 4 * DEC3_PL; B_LO = A_END; B_END = B_LO + 4 * DEC2_PL; C_LO = B_END; C_END = C_LO + 6 * DEC1_PL
SMEM = C_END + 1024 + 128
AB = dict(P1=(B_LO + G * 900 * 4, 1, G * 2 * 29 + 6, 29, 0), P2=(B_LO, 8, G * 25 + 4, 25, 0), P3=(B_LO, 4, G * 20 + 4, 20, 0),
          P4=(B_LO, 1, G * 32 + 16, 32, 0), E4=(B_LO + 4 * (G * 32 + 16) * 16, 2, G * 24 + 8, 24, 15), D1=(C_LO, 4, DEC1_RL, 24, 15),
          D2=(B_LO, 8, DEC2_RL, 28, 31), D3=(A_LO, 16, DEC3_RL, 27, 15), FIN=(A_LO, 32, G * 2 * 12 + 5, 12, 0),
          R3=(C_LO, 8, G * 28 + 1, 28, 0), R4=(C_LO, 16, G * 27 + 1, 27, 0))
RAW1 = AB["E4"][0] + 4 * 2 * AB["E4"][2] * 16

def read_ab(sm, name, planes, g, L, pl0=0):
    off, PH, RL, Q, PAD = AB[name]
    u = sm[off:].view(np.float16)
    out = np.zeros((L, 8 * planes))
    for cp in range(planes):
        for t in range(L):
            pp = t + PAD
            idx = ((pl0 + cp) * PH + pp % PH) * RL + g * Q + pp // PH
            out[t, 8 * cp:8 * cp + 8] = u[idx * 8:idx * 8 + 8]
    return out

def dump(stage):
    buf = torch.zeros(SMEM + 4096, dtype=torch.uint8, device="cuda")
    lib.cm_nwd_mt_debug_dump(C.c_void_p(buf.data_ptr()), stage)
    dem(traces.copy(), monotone_filter_start=900, verbose=False)
    lib.cm_nwd_mt_debug_dump(None, -1)
    return buf.cpu().numpy()

def rep(name, got, want):
    e = np.abs(got - want)
    print("  %-22s max-abs err %.3e (max |want| %.3e) worst at %s" % (name, e.max(), np.abs(want).max(), np.unravel_index(e.argmax(), e.shape)))

dem = NeuralDemixer(path=W, precision="fp16")
only = [int(a) for a in sys.argv[1:]]
for stage in (only or [0, 2, 3, 5, 6, 7, 8, 9, 11, 12, 14, 15, 17]):
    sm = dump(stage)
    print("stage", stage)
    for g in range(G):
        k = keeps[g]
        if stage == 0:
            u = sm[AB["P1"][0]:].view(np.float16)
            for p in (0, 1):
                seq = k["p1"][p::2]
                got = u[(2 * g + p) * 29 * 8:(2 * g + p) * 29 * 8 + len(seq)].astype(np.float64)
                rep("P1 g%d p%d" % (g, p), got, seq)
        if stage == 2: rep("enc1 g%d" % g, read_ab(sm, "D3", 2, g, 387, 2), k["enc1"])
        if stage == 3: rep("P2 g%d" % g, read_ab(sm, "P2", 2, g, 193), h16(E._pool(k["enc1"])))
        if stage == 5: rep("enc2 g%d" % g, read_ab(sm, "D2", 2, g, 162, 2), k["enc2"])
        if stage == 6: rep("enc3 g%d" % g, read_ab(sm, "D1", 4, g, 65, 2), k["enc3"])
        if stage == 7: rep("enc4 g%d" % g, read_ab(sm, "E4", 4, g, 17), k["enc4"])
        if stage == 8: rep("dec1 g%d" % g, read_ab(sm, "D1", 6, g, 65), k["dec1"])
        if stage == 9: rep("dec2 g%d" % g, read_ab(sm, "D2", 4, g, 162), k["dec2"])
        if stage == 11:
            rep("raw3 g%d" % g, read_ab(sm, "R3", 2, g, 193), k["raw3"])
        if stage == 12: rep("dec3 g%d" % g, read_ab(sm, "D3", 4, g, 387), k["dec3"])
        if stage == 14:
            rep("raw4 g%d" % g, read_ab(sm, "R4", 1, g, 402).reshape(804, 4), k["raw4"])
        if stage == 15:
            off, PH, RL, Q, PAD = AB["FIN"]
            u = sm[off:].view(np.float16)
            hp = np.pad(k["h"], ((255, 255), (0, 0)))
            for p in (0, 1):
                xs = hp[p::2]; xs = np.concatenate([xs, np.zeros((800 - len(xs), 4))]).reshape(-1, 8)
                got = np.zeros((353, 8))
                for s in range(353):
                    idx = (s % 32) * RL + (2 * g + p) * 12 + s // 32
                    got[s] = u[idx * 8:idx * 8 + 8]
                rep("fin g%d p%d" % (g, p), got, xs[:353])
        if stage == 17:
            o = sm[C_LO:].view(np.float32)
            got = np.array([o[g * 928 + t + 4 * (t >> 7)] for t in range(900)], dtype=np.float64)
            rep("out g%d" % g, got, k["out"])
