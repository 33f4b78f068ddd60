import os, sys, ctypes as C
os.environ["RESEP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, synth, weights, _lib
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision="bf16")
mix = synth.synth_batch(16, 32000, 2).cuda()
eng = sep._engine
# one intra layer on 64800 rows: resep_layer_fwd(block 0, layer 0)
x = torch.randn(432 * 150, 128, device="cuda")
need = C.c_size_t()
lens = (C.c_int64 * 16)(*[32000] * 16)
eng.lib.resep_workspace_bytes(eng.handle, 16, lens, 2, C.byref(need))
ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
for _ in range(3):
    rc = eng.lib.resep_layer_fwd(eng.handle, 0, 0, x.data_ptr(), 432, 150, ws.data_ptr(), ws.numel(), 2, None)
    assert rc == 0
torch.cuda.synchronize()
buf = (C.c_longlong * 256)()
lib = eng.lib
lib.resep_debug_trace.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.resep_debug_trace(buf))
m = [buf[i] for i in range(64)]; e = [buf[i] for i in range(64, 128)]
t0 = m[0]
print("MMA:", [x - t0 for x in m if x])
print("EPI:", [x - t0 for x in e if x])
u = [buf[i] for i in range(128, 256)]
print("UNIT wait_start, wait_cycles:", [(u[2*i] - t0, u[2*i+1] - u[2*i]) for i in range(64) if u[2*i]])
