"""Clock trace of k_post2_tc (CTA 0: MMA issuer, group A warp 3, group B warp 11) on one 64,800-row layer."""
import os, sys, ctypes as C
os.environ["RESEP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, synth, weights, _lib
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision="bf16")
eng = sep._engine
x = torch.randn(432 * 150, 128, device="cuda")
need = C.c_size_t()
lens = (C.c_int64 * 16)(*[32000] * 16)
eng.lib.resep_workspace_bytes(eng.handle, 16, lens, 2, C.byref(need))
ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
for _ in range(3):
    rc = eng.lib.resep_layer_fwd(eng.handle, 0, 0, x.data_ptr(), 432, 150, ws.data_ptr(), ws.numel(), 2, None)
    assert rc == 0
torch.cuda.synchronize()
buf = (C.c_longlong * 1536)()
lib = eng.lib
lib.resep_debug_trace.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.resep_debug_trace(buf))
t0 = min(buf[r * 512 + 1] for r in range(3) if buf[r * 512 + 1])
print("setup: params staged", buf[1531] - t0, "barriers initialised", buf[1530] - t0, "tmem allocated", buf[1529] - t0)
print("kernel entry", buf[1535] - t0, "setup done", buf[1534] - t0, "producer warp done", buf[1533] - t0, "exit", buf[1532] - t0)
for r, name in enumerate(("MMA", "A", "B")):
    ev = [(buf[r * 512 + 2 * i], buf[r * 512 + 2 * i + 1] - t0) for i in range(256) if buf[r * 512 + 2 * i + 1]]
    print(name, " ".join(f"{t}:{c}" for t, c in ev))
