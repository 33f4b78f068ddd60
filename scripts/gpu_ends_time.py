"""Event-pair times of the HBM-end kernels inside whole config-2 forwards (development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, synth, weights
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision="bf16")
mix = synth.synth_batch(16, 32000, 2).cuda()
for _ in range(3):
    sep.separate_batch(mix)
torch.cuda.synchronize()
prof = sep.profile_kernels(lambda: [sep.separate_batch(mix) for _ in range(10)])
print({k: round(1e3 * v["ms"] / v["launches"], 2) for k, v in prof.items() if not k.startswith(("k_post", "k_qkv", "k_attention"))}, "us per launch")
