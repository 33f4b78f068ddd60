"""Throughput of separate_stream at config 2 against the number of compute lanes (RESEP_LANES, read at import)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, synth, weights
from clearconverse_b200.separation import N_LANES
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision=prec, batch_mode="coupled")
mixes = [synth.synth_batch(16, 32000, 10 + i).cuda() for i in range(64)]
depth = max(2, N_LANES)
for _ in sep.separate_stream((mixes[i % 64] for i in range(8)), depth=depth, device_out=True):
    pass
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in sep.separate_stream((mixes[i % 64] for i in range(60)), depth=depth, device_out=True):
        pass
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 60)
print(f"lanes {N_LANES} depth {depth} {prec}: {best:.4f} ms/step, {16 * 4 / (best / 1e3):.0f} audio-s/s")
