"""The three layer kernels of one 64,800-row intra layer, each 32 times back to back (serialised, no PDL) between one
event pair, and the whole layer 64 times with PDL: the numbers bench.py's roofline block reports (development probe)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, weights
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n_chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 432
code = {"bf16": 2, "fp16": 3}[prec]
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision=prec)
eng = sep._engine; lib = eng.lib
lens = (C.c_int64 * 16)(*[32000] * 16); need = C.c_size_t()
lib.resep_workspace_bytes(eng.handle, 16, lens, code, C.byref(need))
ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
x = torch.randn(n_chunks * 150, 128, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def layer(n):
    for _ in range(n):
        assert lib.resep_layer_fwd(eng.handle, 0, 1, x.data_ptr(), n_chunks, 150, ws.data_ptr(), ws.numel(), code, st) == 0
def timed(fn, n):
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(1e3 * a.elapsed_time(b) / n)
    return sorted(ts)[3]
layer(3); torch.cuda.synchronize()
out = {}
for which, nm in enumerate(("qkv", "attention", "post")):
    def rep(w=which):
        assert lib.resep_layer_kernel_repeat(eng.handle, 0, 1, w, x.data_ptr(), n_chunks, 150, ws.data_ptr(), ws.numel(), code, 32, 0, st) == 0
    rep(); torch.cuda.synchronize()
    out[nm] = round(timed(rep, 32), 2)
    x.normal_()
out["layer_pdl"] = round(timed(lambda: layer(64), 64), 2)
print(prec, n_chunks, "chunks:", out)
