"""One bf16 forward pass of BASELINE config 2 (16 x 4 s), eager launches: the target for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("RESEP_GRAPH", "0")
import torch
from clearconverse_b200 import SepformerSeparation, synth, weights
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision="bf16")
mix = synth.synth_batch(16, 32000, 1).cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    out = sep.separate_batch(mix)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
