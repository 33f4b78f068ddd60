"""Per-kernel time of one config-2 forward in a given precision mode (event-timed, serialised; development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, synth, weights
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision=prec)
mix = synth.synth_batch(16, 32000, 1).cuda()
for _ in range(2): sep.separate_batch(mix)
torch.cuda.synchronize()
prof = sep.profile_kernels(lambda: sep.separate_batch(mix))
tot = sum(v["ms"] for v in prof.values())
print(prec, "total", round(tot, 3), "ms")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:40s} x{v['launches']:3d} {v['ms']:8.3f} ms")
