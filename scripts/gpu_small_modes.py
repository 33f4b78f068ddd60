"""Small bf16 / tf32 / fp32 forwards + the 8f ops, as a quick all-modes smoke (compute-sanitizer is not available on this pool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("RESEP_GRAPH", "0")
import torch
from clearconverse_b200 import SepformerSeparation, synth, weights
sds = weights.random_init_state_dicts(0)
for prec in (sys.argv[1:] or ["bf16"]):
    sep = SepformerSeparation(sds, device="cuda:0", precision=prec, batch_mode="coupled")
    out = sep.separate_batch(synth.synth_batch(2, 2000, 1), peak_normalize=True)
    segs = [synth.synth_mixture(n, 5 + i)[0] for i, n in enumerate([16, 1211, 4000])]
    outs = sep.separate_segments(segs, peak_normalize=True)
    o16 = sep.separate_batch(synth.synth_batch(1, 3001, 2), sample_rate=16000)
    torch.cuda.synchronize()
    print(prec, "ok", float(out.abs().mean()), [tuple(o.shape) for o in outs], tuple(o16.shape))
    sep.close()
