"""Probe: two forwards in flight on two CUDA streams (two separators = two workspaces / graph sets) against one
stream, config 2, device-resident inputs: does batch i+1's intra block fill the GPU while batch i is in its
latency-bound memory block?  (development probe)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, synth, weights
sds = weights.random_init_state_dicts(0)
K = 40
seps = [SepformerSeparation(sds, device="cuda:0", precision="bf16") for _ in range(3)]
mixes = [synth.synth_batch(16, 32000, 1 + i).cuda() for i in range(3)]
streams = [torch.cuda.Stream() for _ in range(3)]
def run(n_streams):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        j = i % n_streams
        with torch.cuda.stream(streams[j]):
            seps[j].separate_batch(mixes[j])
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e3
for n in (1, 2, 3):
    for _ in range(2): run(n)
    print(f"{n} stream(s): {run(n):.4f} ms per forward, {16*4/run(n)*1e3:.0f} audio-s/s")
