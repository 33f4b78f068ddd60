"""In-situ kernel timeline of one graph-replayed bf16 forward (config 2) from torch.profiler / CUPTI: per kernel name
the summed duration, and the summed idle gaps between consecutive kernels (development probe; not a bench number)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from clearconverse_b200 import SepformerSeparation, synth, weights
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision="bf16")
mix = synth.synth_batch(16, 32000, 1).cuda()
for _ in range(5):
    sep.separate_batch(mix)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        sep.separate_batch(mix)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name],
            key=lambda e: e.time_range.start)
# last forward: from the last encoder kernel on
starts = [i for i, e in enumerate(ev) if "k_encoder" in e.name]
ev = ev[starts[-1]:]
t0 = ev[0].time_range.start
tot = ev[-1].time_range.end - t0
per, gap, ovl = collections.OrderedDict(), 0.0, 0.0
prev_end = ev[0].time_range.start
for e in ev:
    n = e.name.split("(")[0].replace("resep::", "").replace("void ", "")
    d = per.setdefault(n, [0, 0.0]); d[0] += 1; d[1] += e.time_range.end - e.time_range.start
    g = e.time_range.start - prev_end
    if g > 0: gap += g
    else: ovl += -g
    prev_end = max(prev_end, e.time_range.end)
print(f"forward span {tot:.1f} us, sum of kernel durations {sum(v[1] for v in per.values()):.1f} us, idle gaps {gap:.1f} us, overlap {ovl:.1f} us")
for n, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n:44s} x{v[0]:3d} {v[1]:8.1f} us")
if len(sys.argv) > 1:
    for e in ev[:int(sys.argv[1])]:
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f} {e.name[:60]}")
