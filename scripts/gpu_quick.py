"""Ad-hoc GPU timing of the forward pass (not the bench contract; used while developing)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, synth, weights

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T = int(sys.argv[3]) if len(sys.argv) > 3 else 32000
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision=prec)
mix = synth.synth_batch(B, T, 2).cuda()
for _ in range(3):
    sep.separate_batch(mix)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    sep.separate_batch(mix)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{prec} B={B} T={T}: {ms:.3f} ms/step, {B*T/8000/(ms/1e3):.0f} audio-s/s")
