"""Throughput / latency of the other BASELINE.json configs on one B200 (bf16 mode unless noted): config 1 shape on
the GPU (1 x 4 s latency), config 3 (1 x 60 s), config 4 (synthetic 1-hour meeting, ~720 s of 0.5-30 s overlap
segments, length-bucketed ragged batches, per-item semantics) and a reduced config-5 sweep.  Writes one JSON
document (committed under profiles/ by hand).  Device-resident inputs, CUDA-event timing, 3 warm-ups."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearconverse_b200 import SepformerSeparation, synth, weights, sharding

dev = torch.device("cuda", 0)
sds = weights.random_init_state_dicts(0)
out = {"gpu": torch.cuda.get_device_name(0), "precision": "bf16", "unit": "audio-s/s"}


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps      # ms


sep = SepformerSeparation(sds, device=dev, precision="bf16", batch_mode="coupled")
sweep = []
for secs in (1, 4, 15, 30):
    for B in (1, 4, 16, 64):
        if B * secs > 1000:
            continue
        mix = synth.synth_batch(B, secs * 8000, seed=7).to(dev)
        ms = timed(lambda: sep.separate_batch(mix), reps=5 if B * secs > 200 else 10)
        sweep.append({"seconds": secs, "batch": B, "ms": round(ms, 3), "audio_s_per_s": round(B * secs / (ms / 1e3), 1)})
        print(sweep[-1], flush=True)
out["config5_sweep_coupled"] = sweep
out["config1_shape_on_gpu"] = next(s for s in sweep if s["seconds"] == 4 and s["batch"] == 1)
mix60 = synth.synth_batch(1, 480000, seed=3).to(dev)
ms = timed(lambda: sep.separate_batch(mix60), reps=10)
out["config3_1x60s"] = {"ms": round(ms, 3), "audio_s_per_s": round(60 / (ms / 1e3), 1)}
print(out["config3_1x60s"], flush=True)
for prec in ("fp16", "tf32", "fp32"):
    s2 = SepformerSeparation(sds, device=dev, precision=prec, batch_mode="coupled")
    mix = synth.synth_batch(16, 32000, seed=2).to(dev)
    ms = timed(lambda: s2.separate_batch(mix), reps=3, warm=2)
    out[f"config2_{prec}"] = {"ms": round(ms, 3), "audio_s_per_s": round(64 / (ms / 1e3), 1)}
    print(prec, out[f"config2_{prec}"], flush=True)
    s2.close()

# config 4: the 1-hour meeting's overlap segments on ONE GPU (rank 0 of 1), ragged batches, per-item semantics
lens = synth.meeting_overlap_segments(720.0, seed=4)
segs = [synth.synth_mixture(n, 5000 + i)[0].to(dev) for i, n in enumerate(lens)]
sep_i = SepformerSeparation(sds, device=dev, precision="bf16", batch_mode="independent")
def run_meeting():
    res, samples = sharding.separate_sharded(segs, sep_i.separate_segments, 0, 1, gather=False)
    return samples
run_meeting(); run_meeting()
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    samples = run_meeting()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
batches, _ = sharding.plan_shards([int(s.numel()) for s in segs], 1)
out["config4_meeting_1gpu"] = {"segments": len(lens), "audio_s": round(sum(lens) / 8000, 1), "batches": len(batches),
                               "wall_ms": round(dt * 1e3, 2), "audio_s_per_s": round(sum(lens) / 8000 / dt, 1),
                               "note": "device-resident segments, host-side bucketing + launches inside the timed region"}
print(out["config4_meeting_1gpu"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs_r2.json", "w"), indent=1)
