"""One intra layer (432 chunks) through resep_layer_fwd a few times: the target of ncu captures of the layer kernels."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, weights
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision=prec)
eng = sep._engine
n_seq = 432
x = torch.randn(n_seq * 150, 128).cuda()
lens = (C.c_int64 * 1)(16 + 8 * (n_seq * 150 + 300)); need = C.c_size_t()
assert eng.lib.resep_workspace_bytes(eng.handle, 1, lens, 2, C.byref(need)) == 0
ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
for _ in range(4):
    rc = eng.lib.resep_layer_fwd(eng.handle, 0, 0, x.data_ptr(), n_seq, 150, ws.data_ptr(), ws.numel(), {"bf16": 2, "fp16": 3}[prec],
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
torch.cuda.synchronize()
print("ok")
