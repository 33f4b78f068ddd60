"""SM clock and board power while one 64,800-row layer runs back to back for ~2 s (development probe)."""
import os, sys, threading, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
from clearconverse_b200 import SepformerSeparation, weights
sep = SepformerSeparation(weights.random_init_state_dicts(0), device="cuda:0", precision="bf16")
eng = sep._engine; lib = eng.lib
lens = (C.c_int64 * 16)(*[32000] * 16); need = C.c_size_t()
lib.resep_workspace_bytes(eng.handle, 16, lens, 2, C.byref(need))
ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
x = torch.randn(432 * 150, 128, device="cuda")
pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False
def sampler():
    while not stop:
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hnd) / 1e3))
        time.sleep(0.02)
th = threading.Thread(target=sampler); th.start()
def run(n):
    for _ in range(n):
        lib.resep_layer_fwd(eng.handle, 0, 1, x.data_ptr(), 432, 150, ws.data_ptr(), ws.numel(), 2, None)
run(20); torch.cuda.synchronize()
for rep in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); a.record(); run(3000); b.record(); torch.cuda.synchronize(); t1 = time.time()
    s = [(c, p) for (t, c, p) in samples if t0 + 0.05 < t < t1]
    print(f"rep {rep}: {a.elapsed_time(b) / 3000 * 1e3:.2f} us/layer; sm MHz {sorted(c for c, _ in s)[len(s)//2] if s else None}; W max {max((p for _, p in s), default=None)}")
stop = True; th.join()
