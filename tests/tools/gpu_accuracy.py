"""Accuracy of every precision mode against the oracle for several weight seeds -> gpurun_out/accuracy_r2.json
(copied to profiles/r2_accuracy.json): max-abs, est-vs-oracle-est SI-SNR, and per (item, speaker) the oracle's SI-SNR
against the mixture with the SI-SNR delta next to it -- the delta is only meaningful where the reference SI-SNR is
well-conditioned (tests/test_gpu_parity.py::si_snr_delta)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from clearconverse_b200 import SepformerSeparation
from clearconverse_b200.synth import synth_batch
from oracle.resepformer_oracle import OracleSepformerSeparation
from clearconverse_b200.metrics import si_snr_db
torch.set_num_threads(os.cpu_count())
cases = [(2, 2000, 2), (1, 32000, 1), (3, 9000, 5), (8, 32000, 2)]
out = {"note": "coupled batches, synthetic 8 kHz mixtures, random-init weights per seed; delta_db_well_conditioned = max |SI-SNR(est,mix) - "
               "SI-SNR(oracle_est,mix)| over (item, speaker) pairs with oracle SI-SNR >= -20 dB", "rows": []}
from clearconverse_b200 import weights as _w
def _fb_oracle(seed):
    base = OracleSepformerSeparation(seed=seed).component_state_dicts()
    m = OracleSepformerSeparation(seed=None, distinct_blocks=False)
    fb = _w.filterbank_init_state_dicts(base=base)
    for k in ("encoder", "masknet", "decoder"):
        m.mods[k].load_state_dict(fb[k])
    return m
# weight sets: random init of seeds 0..3, and the well-conditioned filterbank set over the masknets of seeds 0 and 3
for seed, oracle in [(s_, OracleSepformerSeparation(seed=s_)) for s_ in (0, 1, 2, 3)] + [(f"filterbank/{s_}", _fb_oracle(s_)) for s_ in (0, 3)]:
    wants = [(synth_batch(*c), None) for c in cases]
    wants = [(x, oracle.separate_batch(x)) for x, _ in wants]
    for prec in ("fp32", "tf32", "fp16", "bf16"):
        sep = SepformerSeparation(oracle.component_state_dicts(), device="cuda:0", precision=prec)
        for c, (x, w) in zip(cases, wants):
            g = sep.separate_batch(x).cpu()
            R = si_snr_db(w.permute(0, 2, 1), x[:, None, :]).flatten()
            dl = (si_snr_db(g.permute(0, 2, 1), x[:, None, :]).flatten() - R).abs()
            ok = R >= -20.0
            row = {"weight_seed": seed, "precision": prec, "B": c[0], "T": c[1], "max_abs": float((g - w).abs().max()),
                   "est_vs_oracle_est_si_snr_db_min": float(si_snr_db(g.permute(0, 2, 1), w.permute(0, 2, 1)).min()),
                   "delta_db_well_conditioned": float(dl[ok].max()) if ok.any() else None,
                   "delta_db_all": float(dl.max()), "oracle_si_snr_vs_mix_db_min": float(R.min())}
            out["rows"].append(row)
            print(row, flush=True)
        sep.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "accuracy_r2.json"), "w"), indent=1)
