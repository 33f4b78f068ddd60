"""bf16-mode accuracy against the oracle for several weight seeds (development probe): max-abs, est-vs-est SI-SNR,
and per (item, speaker) the oracle's SI-SNR against the mixture with the SI-SNR delta next to it -- the delta is only
meaningful where the reference SI-SNR is well-conditioned (see tests/test_gpu_parity.py::si_snr_delta)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from clearconverse_b200 import SepformerSeparation
from clearconverse_b200.synth import synth_batch
from oracle.resepformer_oracle import OracleSepformerSeparation
from test_gpu_parity import si_snr_db
torch.set_num_threads(os.cpu_count())
for seed in (0, 1, 2, 3):
    oracle = OracleSepformerSeparation(seed=seed)
    sep = SepformerSeparation(oracle.component_state_dicts(), device="cuda:0", precision="bf16")
    for x in (synth_batch(2, 2000, 2), synth_batch(1, 32000, 1), synth_batch(3, 9000, 5), synth_batch(8, 32000, 2)):
        w = oracle.separate_batch(x)
        g = sep.separate_batch(x).cpu()
        R = si_snr_db(w.permute(0, 2, 1), x[:, None, :]).flatten()
        dl = (si_snr_db(g.permute(0, 2, 1), x[:, None, :]).flatten() - R).abs()
        print(f"seed {seed} B={x.shape[0]} T={x.shape[1]} maxabs {(g-w).abs().max():.1e} est-vs-est {si_snr_db(g.permute(0,2,1), w.permute(0,2,1)).min():.1f} dB |",
              " ".join(f"{r:.0f}:{d:.3f}" for r, d in zip(R.tolist(), dl.tolist())), flush=True)
    sep.close()
