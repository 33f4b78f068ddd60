"""Accuracy of the bf16 mode's weight-operand variants against the oracle (development probe)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
from clearconverse_b200 import SepformerSeparation
from clearconverse_b200.synth import synth_batch
from oracle.resepformer_oracle import OracleSepformerSeparation
from clearconverse_b200.metrics import si_snr_db, si_snr_delta
torch.set_num_threads(os.cpu_count())
oracle = OracleSepformerSeparation(seed=0)
sds = oracle.component_state_dicts()
cases = [synth_batch(2, 2000, 2), synth_batch(1, 32000, 1), synth_batch(3, 9000, 5), synth_batch(16, 32000, 2)]
wants = [oracle.separate_batch(x) for x in cases]
# linear hook, precision codes 2,3,4
for mode in ("bf16", "bf16x2", "mixed"):
    os.environ["RESEP_W16"] = mode
    sep = SepformerSeparation(sds, device="cuda:0", precision="bf16")
    row = []
    for x, w in zip(cases, wants):
        g = sep.separate_batch(x).cpu()
        row.append(f"{(g-w).abs().max():.1e} {si_snr_db(g.permute(0,2,1), w.permute(0,2,1)).min():.1f}dB d={si_snr_delta(g,w,x)[0]:.4f}")
    print(f"W16={mode:7s}", " | ".join(row), flush=True)
    sep.close()
sep = SepformerSeparation(sds, device="cuda:0", precision="tf32")
row = []
for x, w in zip(cases, wants):
    g = sep.separate_batch(x).cpu()
    row.append(f"{(g-w).abs().max():.1e} {si_snr_db(g.permute(0,2,1), w.permute(0,2,1)).min():.1f}dB d={si_snr_delta(g,w,x)[0]:.4f}")
print("tf32(splitW)", " | ".join(row), flush=True)
eng = sep._engine
g = torch.Generator().manual_seed(3)
A = torch.randn(1000, 128, generator=g).cuda(); W = (torch.randn(384, 128, generator=g) / 128 ** 0.5).cuda(); b = torch.randn(384, generator=g).cuda()
ref = A.double() @ W.double().T + b.double()
for prec in (0, 1, 2, 3, 4):
    o = torch.empty(1000, 384, device="cuda")
    rc = eng.lib.resep_linear_fwd(eng.handle, A.data_ptr(), W.data_ptr(), b.data_ptr(), o.data_ptr(), 1000, 384, 128, 0, prec, None)
    print("linear prec", prec, "rc", rc, "max err", (o.double() - ref).abs().max().item())
