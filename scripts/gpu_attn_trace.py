"""Development probe: clock trace of k_attn_tc's MMA warp and first softmax warp (CTA 0) for one intra-layer launch.
Needs a build with EXTRA=-DRESEP_TRACE_BUILD and RESEP_TRACE=1 in the environment."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, weights
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision="bf16")
eng = sep._engine
n_seq = 432
x = torch.randn(n_seq * 150, 128).cuda()
lens = (C.c_int64 * 1)(16 + 8 * (n_seq * 150 + 300)); need = C.c_size_t()
assert eng.lib.resep_workspace_bytes(eng.handle, 1, lens, 2, C.byref(need)) == 0
ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
for _ in range(2):
    rc = eng.lib.resep_layer_fwd(eng.handle, 0, 0, x.data_ptr(), n_seq, 150, ws.data_ptr(), ws.numel(), 2, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
torch.cuda.synchronize()
buf = (C.c_longlong * 2048)()
eng.lib.resep_debug_attn_trace.argtypes = [C.c_void_p]
assert eng.lib.resep_debug_attn_trace(buf) == 0
for role, name in ((0, "mma"), (1, "softmax")):
    ev = [(buf[role * 1024 + 2 * i], buf[role * 1024 + 2 * i + 1]) for i in range(512) if buf[role * 1024 + 2 * i]]
    if not ev:
        continue
    t0 = ev[0][1]
    print(name, "events", len(ev))
    prev = t0
    for tag, t in ev[:140]:
        print(f"  {name} {tag:5d} +{t - prev:6d}  @{t - t0}")
        prev = t
