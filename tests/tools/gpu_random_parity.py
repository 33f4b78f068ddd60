"""One-off robustness sweep: random ragged batches in every precision mode against per-item oracle calls."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from clearconverse_b200 import SepformerSeparation
from clearconverse_b200.synth import synth_mixture
from oracle.resepformer_oracle import OracleSepformerSeparation
from clearconverse_b200.metrics import si_snr_db
torch.set_num_threads(os.cpu_count())
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
oracle = OracleSepformerSeparation(seed=0)
seps = {p: SepformerSeparation(oracle.component_state_dicts(), device="cuda:0", precision=p, batch_mode="independent") for p in ("fp32", "tf32", "fp16", "bf16")}
worst = {p: 0.0 for p in seps}
for trial in range(6):
    lens = [rng.choice([16, 17, rng.randint(16, 3000), rng.randint(16, 20000), 1200 * rng.randint(1, 6) + rng.choice([0, 7, 8, 15])]) for _ in range(rng.randint(1, 6))]
    segs = [synth_mixture(n, 1000 * trial + i)[0] for i, n in enumerate(lens)]
    wants = [oracle.separate_batch(s[None])[0] for s in segs]
    for p, sep in seps.items():
        outs = sep.separate_segments(segs, peak_normalize=False)
        for n, o, w in zip(lens, outs, wants):
            assert o.shape == w.shape, (p, n)
            d = (o.cpu() - w).abs().max().item()
            worst[p] = max(worst[p], d)
            assert torch.isfinite(o).all(), (p, n)
    print(trial, lens, {p: f"{v:.2e}" for p, v in worst.items()}, flush=True)
assert worst["fp32"] < 1e-4 and worst["tf32"] < 1e-3 and worst["fp16"] < 1e-3 and worst["bf16"] < 5e-2
print("ok", worst)
