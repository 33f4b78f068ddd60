"""One transformer layer through the C ABI (resep_layer_fwd): bf16 tensor-core path vs the fp32 FMA path on
the device, for several row counts, plus per-kernel timing of a 64,800-row layer (development probe)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, weights
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision="bf16")
eng = sep._engine
lib = eng.lib
lens = (C.c_int64 * 16)(*[32000] * 16)
need = C.c_size_t()
lib.resep_workspace_bytes(eng.handle, 16, lens, 2, C.byref(need))
ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
g = torch.Generator().manual_seed(5)
for n_seq in (1, 2, 3, 27, 149, 432):
    x = (torch.randn(n_seq * 150, 128, generator=g) * 2 + 0.3).cuda()
    ref = x.clone(); got = x.clone()
    assert lib.resep_layer_fwd(eng.handle, 0, 1, ref.data_ptr(), n_seq, 150, ws.data_ptr(), ws.numel(), 0, None) == 0
    rc = lib.resep_layer_fwd(eng.handle, 0, 1, got.data_ptr(), n_seq, 150, ws.data_ptr(), ws.numel(), 2, None)
    assert rc == 0, lib.resep_last_error(eng.handle)
    torch.cuda.synchronize()
    err = (got - ref).abs()
    print(f"n_seq={n_seq:4d} rows={n_seq*150:6d} max|bf16-fp32|={err.max().item():.3e} mean={err.mean().item():.3e} "
          f"ref max={ref.abs().max().item():.2f} finite={bool(torch.isfinite(got).all())}", flush=True)
    if err.max().item() > 0.2:
        bad = (err > 0.2).nonzero()
        print("  bad rows:", sorted(set((bad[:, 0] // 128).tolist()))[:20], "first bad", bad[:5].tolist())
x = torch.randn(432 * 150, 128, generator=g).cuda()
def run():
    for _ in range(10):
        lib.resep_layer_fwd(eng.handle, 0, 1, x.data_ptr(), 432, 150, ws.data_ptr(), ws.numel(), 2, None)
run(); torch.cuda.synchronize()
prof = sep.profile_kernels(run)
print({k: round(1e3 * v["ms"] / v["launches"], 2) for k, v in prof.items()}, "us per launch")
