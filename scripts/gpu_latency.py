"""Latency of one B = 1 call (the reference's call pattern, api.py:1077) per precision mode and segment length."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, synth, weights
sds = weights.random_init_state_dicts(0)
for prec in ("bf16", "tf32", "fp32"):
    sep = SepformerSeparation(sds, device="cuda:0", precision=prec)
    row = []
    for secs in (1, 4, 15):
        mix = synth.synth_batch(1, 8000 * secs, secs).cuda()
        for _ in range(5): sep.separate_batch(mix)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(30): out = sep.separate_batch(mix)
        torch.cuda.synchronize()
        row.append(f"{secs} s: {(time.perf_counter() - t0) / 30 * 1e3:.3f} ms")
    print(prec, "graphs" if os.environ.get("RESEP_GRAPH") != "0" else "eager", " | ".join(row), flush=True)
    sep.close()
