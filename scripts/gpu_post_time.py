"""Time one 64,800-row transformer layer's kernels (development probe; RESEP_DBG variants are not correct results)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, weights
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision="bf16")
eng = sep._engine; lib = eng.lib
lens = (C.c_int64 * 16)(*[32000] * 16); need = C.c_size_t()
lib.resep_workspace_bytes(eng.handle, 16, lens, 2, C.byref(need))
ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
x = torch.randn(432 * 150, 128, device="cuda")
def run():
    for _ in range(10):
        lib.resep_layer_fwd(eng.handle, 0, 1, x.data_ptr(), 432, 150, ws.data_ptr(), ws.numel(), 2, None)
run(); torch.cuda.synchronize()
prof = sep.profile_kernels(run)
print("DBG", os.environ.get("RESEP_DBG"), {k: round(1e3 * v["ms"] / v["launches"], 2) for k, v in prof.items()}, "us per launch")
