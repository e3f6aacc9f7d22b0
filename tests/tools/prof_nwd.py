"""Profiling workload (not a test): C2 NWD forward on the fp16 multi-trace tensor-core path, for ncu captures."""
import sys
import torch
sys.path.insert(0, ".")
from bench import synth_traces
from circuitmap_b200 import NeuralDemixer
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dem = NeuralDemixer(path="tests/golden/nwd_ie_ChroME2f_weights.npz", precision=sys.argv[2] if len(sys.argv) > 2 else "fp16")
x = torch.from_numpy(synth_traces(K, 0)).float().cuda()
for _ in range(4):
    o = dem.forward_device(x)
torch.cuda.synchronize()
print("nwd ok", float(o.abs().max()))
