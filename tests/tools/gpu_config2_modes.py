"""Max-abs of the 1e-3-contract modes (fp16, tf32) against the oracle at the full config-2 batch, both batch semantics."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from clearconverse_b200 import SepformerSeparation
from clearconverse_b200.synth import synth_batch
from oracle.resepformer_oracle import OracleSepformerSeparation
torch.set_num_threads(os.cpu_count())
for wseed in (0, 1):
    oracle = OracleSepformerSeparation(seed=wseed)
    sds = oracle.component_state_dicts()
    mix = synth_batch(16, 32000, 2)
    want = oracle.separate_batch(mix)
    for mode in ("coupled", "independent"):
        w = want if mode == "coupled" else torch.cat([oracle.separate_batch(mix[i:i + 1]) for i in range(16)])
        for prec in ("fp16", "tf32"):
            with SepformerSeparation(sds, device="cuda:0", precision=prec, batch_mode=mode) as sep:
                got = sep.separate_batch(mix).cpu()
            print(wseed, mode, prec, f"{(got - w).abs().max().item():.3e}", flush=True)
