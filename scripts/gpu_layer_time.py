"""Device time of one 64,800-row transformer layer (qkv2 + attention + post2 back to back, PDL on), events around 200 calls."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, weights
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision=prec)
eng = sep._engine; lib = eng.lib
code = {"bf16": 2, "fp16": 3}[prec]
lens = (C.c_int64 * 16)(*[32000] * 16); need = C.c_size_t()
lib.resep_workspace_bytes(eng.handle, 16, lens, code, C.byref(need))
ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
x = torch.randn(432 * 150, 128, device="cuda")
def run(n):
    for _ in range(n):
        rc = lib.resep_layer_fwd(eng.handle, 0, 1, x.data_ptr(), 432, 150, ws.data_ptr(), ws.numel(), code, None)
        assert rc == 0
run(20); torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(200); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 200 * 1e3)
print(f"{prec}: {best:.2f} us per layer (200 layers back to back, best of 5)")
