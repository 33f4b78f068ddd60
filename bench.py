#!/usr/bin/env python
"""Benchmark of the RE-SepFormer overlap-separation path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: BASELINE.json
configs[1], 16 x 4 s synthetic 8 kHz two-speaker mixtures through ``separate_batch`` in the
bf16 mode on one B200.  With N > 1 every rank runs that same per-GPU batch (independent
segments shard with no data-path collective -> weak scaling) and ``value`` is the whole-job
audio-seconds per second with the step time taken as the max over ranks.

Prints ONE JSON line (rank 0).  ``value`` is measured with the inputs resident in HBM; ``e2e``
goes through the public Python API with pinned HOST buffers, H2D and D2H inside the timed
region.  ``roofline`` is the dominant kernel of the step (CUDA-event timing per kernel on the
launch stream, taken in a separate pass); ``cpu_baseline`` is the oracle (the only runnable
form of the reference's CPU path: speechbrain is not installable here) on the host cores.

Side blocks of the same line (none of them changes ``value``):
  ``sustained``   the same step repeated for >= 3 s with clocks and power sampled throughout;
  ``meeting``     BASELINE configs[3]: the overlap segments of a synthetic 1-hour meeting (720 s, 0.5-30 s,
                  log-uniform, seed 4) through ``sharding.separate_sharded`` -- planning, H2D, kernels, D2H and the
                  host gather inside the timed region; the SAME segment list at every N (strong scaling);
  ``long_split``  (N > 1) one 60 s recording split by chunks across the ranks: bit-identity flag and time;
  ``latency``     one B = 1 call of 4 s (the reference's call pattern), repeated shape vs. a stream of distinct lengths.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "separated_audio_seconds_per_second"
UNIT = "audio-s/s"
SAMPLE_RATE = 8000
BATCH, SECONDS = 16, 4
T = SECONDS * SAMPLE_RATE
DEPTH = 4          # forwards in flight in the pipelined driver (separate_stream: one CUDA stream, workspace and graph set per lane)
WORKLOAD = "resepformer-wsj02mix random-init, batch 16 x 4 s synthetic 8 kHz 2-speaker mixtures (BASELINE configs[1])"


def peaks():
    fb = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_src": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        d["_src"] = "measured"
        return d
    except Exception:
        return fb


# ------------------------------------------------------------------------------------- model shapes
def shapes(batch, t, coupled=True):
    L = (t - 16) // 8 + 1
    S = L // 150 + 1
    chunks = batch * S
    M = chunks * 150
    mem_seqs = [chunks] if coupled else [S] * batch
    return L, S, chunks, M, mem_seqs


def kernel_work(name: str, batch: int, t: int):
    """(bound, algorithmic FLOPs or bytes summed over ALL launches of this kernel in one step).
    Per-unit figures are SURVEY.md section 8(d): multiply-add = 2 FLOPs; per token-layer QKV 98,304,
    attention 76,800, out-proj 32,768, FFN 524,288."""
    L, S, chunks, M, mem_seqs = shapes(batch, t)
    rows = 16 * M + 8 * chunks                         # token-layers: 16 intra layers + 8 memory layers
    if name.endswith("(small)"):                       # the memory transformer's launches of the CTA-pair kernels
        rows, name = 8 * chunks, name[:-7]
    elif name.startswith("k_post2_tc") or name.startswith("k_qkv2_tc"):
        rows = 16 * M
    att_intra = 16 * chunks * 8 * 4 * 150 * 150 * 16
    att = att_intra + 8 * sum(4 * n * n * 128 for n in mem_seqs)
    if name.startswith("k_gemm_tc<bf16,bf16"):        # QKV + FFN1
        return "tensor", rows * (98_304 + 262_144)
    if name.startswith("k_gemm_tc<bf16,resid"):       # out-proj + FFN2
        return "tensor", rows * (32_768 + 262_144)
    if name.startswith("k_gemm_tc<bf16,f32"):         # output_fc
        return "tensor", M * 65_536
    if name.startswith("k_post"):                      # fused out-proj + LN2 + FFN1 + ReLU + FFN2 (+ residuals); k_post2_tc = CTA-pair version
        return "tensor", rows * (32_768 + 524_288)
    if name.startswith("k_qkv"):                       # fused LN1 + in-projection: 512 B in + 768 B out per row, HBM-bound
        return "hbm", rows * (512 + 768)
    if name.startswith("k_attention_bf16_tma") or name.startswith("k_attn_tc"):   # every intra chunk
        return "tensor", att_intra
    if name.startswith("k_attention_bf16_short"):      # sequences <= 160 rows (every intra chunk)
        return "tensor", att_intra + sum(8 * 4 * n * n * 128 for n in mem_seqs if n <= 160)
    if name.startswith("k_attention"):
        return "tensor", att - att_intra - sum(8 * 4 * n * n * 128 for n in mem_seqs if n <= 160) if "bf16" in name else att
    if name.startswith("k_layernorm"):
        return "hbm", 2 * rows * (512 + 256)
    if name.startswith("k_block_epilogue"):
        return "hbm", (2 * M + chunks) * 128 * 4 * 4
    if name.startswith("k_maskdec"):                   # fused output_fc + mask + decoder: PReLU bf16 + features in, samples out
        return "hbm", M * (256 + 512) + batch * t * 8
    if name.startswith("k_decoder"):
        return "hbm", batch * L * (256 + 128) * 4 + batch * t * 8
    if name.startswith("k_encoder"):
        return "hbm", batch * t * 4 + M * 512
    return None, None


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.power, self.reasons, self.max_mhz = [], [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.dev) / 1e3)
                except Exception:
                    pass
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join(1.0)

    def summary(self):
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": (max(self.power) if self.power else None)}


def bench_config(world: int) -> dict:
    """The ``config`` object of the JSON line -- the SAME dict in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH, "seconds_per_item": SECONDS, "sample_rate": SAMPLE_RATE,
            "batch_mode": "coupled", "parallelism": f"replicated x{world}, no collective",
            "l2": "inputs rotated over 64 device-resident batches (128 MiB > the 126 MB L2)"}


# ------------------------------------------------------------------------------------- CPU arm
def make_oracle(sds):
    import torch
    from oracle.resepformer_oracle import OracleSepformerSeparation
    m = OracleSepformerSeparation(seed=None, distinct_blocks=False)
    for k in ("encoder", "masknet", "decoder"):
        m.mods[k].load_state_dict(sds[k])
    return m


def time_oracle(sds, items: int, reps: int, warmup: int = 1):
    """Oracle (fp32 eager PyTorch, all host threads) on `items` of the 16 mixtures per repetition."""
    import torch
    from clearconverse_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = make_oracle(sds)
    mix = synth.synth_batch(items, T, seed=2)
    times = []
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        m.separate_batch(mix)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return items * SECONDS / statistics.median(times), times, torch.get_num_threads()


def time_oracle_other_configs(sds):
    """CPU figures for BASELINE.md section 5 (opt-in, --cpu-other-configs): config 1 (1 x 4 s), config 3 (1 x 60 s) and a
    bounded sample of config 4 as the product runs it (one B = 1 call per overlap segment, api.py:1073-1077)."""
    import torch
    from clearconverse_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    m = make_oracle(sds)
    out = {"cores": torch.get_num_threads()}
    def med(fn, reps):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return statistics.median(ts)
    x1 = synth.synth_batch(1, 4 * SAMPLE_RATE, seed=1)
    out["config1_1x4s_audio_s_per_s"] = 4.0 / med(lambda: m.separate_batch(x1), 5)
    x3 = synth.synth_batch(1, 60 * SAMPLE_RATE, seed=3)
    out["config3_1x60s_audio_s_per_s"] = 60.0 / med(lambda: m.separate_batch(x3), 3)
    lens = synth.meeting_overlap_segments(720.0, seed=4)[:12]
    segs = [synth.synth_mixture(n, 5000 + i)[0].unsqueeze(0) for i, n in enumerate(lens)]
    t = med(lambda: [m.separate_batch(sg) for sg in segs], 1)
    out["config4_b1_loop_audio_s_per_s"] = sum(lens) / SAMPLE_RATE / t
    out["config4_sample"] = f"the first {len(lens)} of the meeting's segments ({sum(lens) / SAMPLE_RATE:.0f} s of audio), one B=1 call each"
    return out


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path.  speechbrain cannot be
    installed here, so this is the oracle port on the host cores; rank 0 alone runs it."""
    if rank != 0:
        return
    from clearconverse_b200 import weights
    sds = weights.random_init_state_dicts(0)
    items = BATCH          # the whole configs[1] batch per step (coupled semantics depend on the batch composition)
    import torch
    from clearconverse_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = make_oracle(sds)
    mix = synth.synth_batch(items, T, seed=2)
    for _ in range(max(1, min(args.warmup, 3))):
        m.separate_batch(mix)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.separate_batch(mix)
    dt = time.perf_counter() - t0
    value = args.steps * items * SECONDS / dt
    sample = f"{items} of the {BATCH} mixtures (B={items} x {SECONDS} s) per step, oracle fp32 eager, {torch.get_num_threads()} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-other-configs", action="store_true",
                    help="also time the oracle on configs 1, 3 and a sample of config 4 (adds about a minute of CPU work)")
    ap.add_argument("--no-side-blocks", action="store_true",
                    help="skip the sustained / meeting / long_split / latency / cpu_baseline blocks (the ncu passes under "
                         "profiles/ use this: the timed step and its kernels are the same, the run is ~100x shorter)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from clearconverse_b200 import SepformerSeparation, synth, weights

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    sds = weights.random_init_state_dicts(0)
    sep = SepformerSeparation(sds, device=dev, precision=args.precision, batch_mode="coupled")
    host_mix = synth.synth_batch(BATCH, T, seed=2 + rank).pin_memory()
    mix = host_mix.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- the layer kernels alone, back to back (the roofline block's durations).  Taken FIRST, at burst clocks:
    # the burst bf16 peak they are compared with was measured the same way, and later sections (the >= 3 s sustained pass)
    # leave the board at its power cap (same kernel, same method: 35.4 us before / 37.0 us after that pass)
    def layer_kernel_us():
        """The three kernels of one intra layer at the bench shape, each launched 32 times back to back between two
        events (no per-launch event records, no host gaps; no PDL: a launch includes its own set-up and tail), and the
        whole layer 64 times back to back with PDL as in the product path.  Median of 5."""
        import ctypes as C
        from clearconverse_b200._lib import PRECISIONS
        eng = sep._engine
        code = PRECISIONS[args.precision]
        n_chunks = shapes(BATCH, T)[2]
        lens = (C.c_int64 * BATCH)(*[T] * BATCH)
        need = C.c_size_t()
        eng.lib.resep_workspace_bytes(eng.handle, BATCH, lens, code, C.byref(need))
        wsb = torch.empty(need.value, dtype=torch.uint8, device=dev)
        x = torch.randn(n_chunks * 150, 128, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        def layer(n):
            for _ in range(n):
                assert eng.lib.resep_layer_fwd(eng.handle, 0, 1, x.data_ptr(), n_chunks, 150, wsb.data_ptr(), wsb.numel(), code, st) == 0
        def timed(fn, n):
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(1e3 * a.elapsed_time(b) / n)
            return sorted(ts)[2]
        layer(3); torch.cuda.synchronize()
        out = {}
        for which, nm in enumerate(("qkv", "attention", "post")):
            def rep(w=which):
                assert eng.lib.resep_layer_kernel_repeat(eng.handle, 0, 1, w, x.data_ptr(), n_chunks, 150, wsb.data_ptr(), wsb.numel(), code, 32, 0, st) == 0
            rep(); torch.cuda.synchronize()
            out[nm] = timed(rep, 32)
            x.normal_()
        out["layer_pdl"] = timed(lambda: layer(64), 64)
        return out
    lk = None
    if args.precision in ("bf16", "fp16") and not args.no_side_blocks:      # (--no-side-blocks: profiler passes want one plain forward)
        try:
            lk = layer_kernel_us()
        except Exception as e:   # the measurement aid must never take the bench line down
            lk = {"error": repr(e)}

    # ---------------- device-resident timing, one forward at a time: K steps, each bracketed by events, L2 flushed in
    # between (the latency view; the per-kernel roofline below refers to this mode)
    for _ in range(args.warmup):
        sep.separate_batch(mix)
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        flush.zero_()
        a.record()
        sep.separate_batch(mix)
        b.record()
    barrier()
    single_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(single_ms, op=dist.ReduceOp.MAX)
    audio_s = world * BATCH * SECONDS * args.steps
    single_value = audio_s / (single_ms.item() / 1e3)

    # ---------------- device-resident throughput (`value`): the same K forwards through the pipelined driver, four in
    # flight on four CUDA streams (a forward's memory transformer is 24 latency-bound launches on 4-56 CTAs and every
    # persistent layer kernel ends in a partly filled last round: the other forwards' kernels fill those SMs).  Inputs are rotated over 64 device-resident batches
    # (128 MiB > the 126 MB L2) instead of flushing L2, which would serialise the lanes.
    NROT = 64
    dev_mixes = [torch.roll(mix, i, 0).contiguous() for i in range(NROT)]
    for _ in sep.separate_stream((dev_mixes[i % NROT] for i in range(max(args.warmup, 3 * DEPTH))), depth=DEPTH, device_out=True):   # every lane: eager, capture, first replay
        pass
    barrier()
    l0 = sep.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        t_wall0 = time.perf_counter()
        ev0.record()
        n_done = 0
        for _ in sep.separate_stream((dev_mixes[i % NROT] for i in range(args.steps)), depth=DEPTH, device_out=True):
            n_done += 1                                    # (each result is synchronised before it is yielded)
        ev1.record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    assert n_done == args.steps
    launches = sep.launch_count() - l0
    total_ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = total_ms.item() / 1e3
    value = audio_s / total_s

    # ---------------- end to end: pinned host in, pinned host out, through the public API.  Every step copies that
    # step's input H2D and its result D2H inside the timed region; `separate_stream` (the batched driver) overlaps
    # the copies of neighbouring steps with the kernels.  The serial form (one blocking separate_batch + copy per
    # step, what a single api.py call does) is reported next to it.
    host_outs = [torch.empty(BATCH, T, 2, dtype=torch.float32).pin_memory() for _ in range(DEPTH + 1)]
    host_ins = [host_mix, host_mix.clone().pin_memory()]
    for _ in sep.separate_stream((host_ins[i & 1] for i in range(3 * DEPTH)), host_outs, depth=DEPTH):
        pass
    barrier()
    t0 = time.perf_counter()
    n_out = 0
    for out in sep.separate_stream((host_ins[i & 1] for i in range(args.steps)), host_outs, depth=DEPTH):
        n_out += 1
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    assert n_out == args.steps
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = audio_s / e2e_s.item()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        est = sep.separate_batch(host_mix)                 # H2D of the step's input inside
        host_outs[0].copy_(est, non_blocking=True)         # D2H of the step's result
        torch.cuda.current_stream().synchronize()
    e2e_serial_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_serial_s, op=dist.ReduceOp.MAX)
    e2e_serial_value = audio_s / e2e_serial_s.item()

    # ---------------- host <-> device copy rates with EVERY rank copying at the same time (the e2e path's only shared
    # resource: PCIe root complexes / host memory; there is no data-path collective).  Compared across N this names
    # what the e2e scaling loses.
    copy_probe = None
    if not args.no_side_blocks:
        barrier()
        cp_s = torch.cuda.Stream(dev)
        n_cp = 200
        with torch.cuda.stream(cp_s):
            ea, eb, ec = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            ea.record(cp_s)
            for i in range(n_cp):
                mix.copy_(host_ins[i & 1], non_blocking=True)
            eb.record(cp_s)
            for i in range(n_cp):
                host_outs[i % 3].view(-1)[:BATCH * T * 2].copy_(flush[:BATCH * T * 2 * 4].view(torch.float32), non_blocking=True)
            ec.record(cp_s)
        cp_s.synchronize()
        t = torch.tensor([ea.elapsed_time(eb), eb.elapsed_time(ec)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        copy_probe = {"h2d_gbs_per_gpu": n_cp * BATCH * T * 4 / (t[0].item() / 1e3) / 1e9,
                      "d2h_gbs_per_gpu": n_cp * BATCH * T * 2 * 4 / (t[1].item() / 1e3) / 1e9,
                      "note": f"{n_cp} x the step's 2.0 MB input H2D, then {n_cp} x its 4.1 MB result D2H, pinned host memory, all {world} ranks at once (slowest rank)",
                      "needed_gbs_per_gpu_at_value": (BATCH * T * 4 + BATCH * T * 2 * 4) / (total_s / args.steps) / 1e9}

    # ---------------- sustained: the same step through the same driver for >= 3 s (clocks and power sampled throughout)
    sust_steps = max(args.steps, int(3.2 / max(total_s / args.steps, 1e-6))) if not args.no_side_blocks else args.steps
    barrier()
    ev0s, ev1s = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks_s:
        ev0s.record()
        for _ in sep.separate_stream((dev_mixes[i % NROT] for i in range(sust_steps)), depth=DEPTH, device_out=True):
            pass
        ev1s.record()
        barrier()
    sust_ms = torch.tensor([ev0s.elapsed_time(ev1s)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(sust_ms, op=dist.ReduceOp.MAX)
    cs = clocks_s.summary()
    sustained = {"value": world * BATCH * SECONDS * sust_steps / (sust_ms.item() / 1e3), "unit": UNIT, "steps": sust_steps,
                 "seconds": sust_ms.item() / 1e3, "sm_mhz_median": cs["sm_mhz"], "power_w_max": cs["power_w_max"],
                 "reasons": cs["reasons"]}

    # ---------------- BASELINE configs[3]: the synthetic 1-hour meeting, sharded over the N GPUs (strong scaling)
    meeting = run_meeting(sds, args.precision, rank, world, dev, dist, barrier) if not args.no_side_blocks else None
    long_split = run_long_split(sep, rank, world, dev, dist) if world > 1 and not args.no_side_blocks else None

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- latency of the reference's own call pattern (api.py:1073-1077): B = 1, one call at a time
    def one_call_ms(mixes, reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(reps):
            sep.separate_batch(mixes[i % len(mixes)])
            torch.cuda.current_stream().synchronize()
        return 1e3 * (time.perf_counter() - t0) / reps
    latency = None
    if not args.no_side_blocks:
        same = [synth.synth_batch(1, T, 5).to(dev)]
        one_call_ms(same, 5)
        distinct = [synth.synth_batch(1, T - 801 * i, 60 + i).to(dev) for i in range(32)]        # 4 s ... 0.9 s, every length new
        latency = {"b1_4s_repeated_shape_ms": one_call_ms(same, 50),
                   "b1_first_sighting_of_each_length_ms": one_call_ms(distinct, 32),              # plan upload + eager launches
                   "b1_distinct_lengths_seen_before_ms": one_call_ms(distinct, 64),               # plans cached; graphs after the 2nd sighting
                   "note": "blocking separate_batch on a device-resident [1,T] mixture, host wall clock per call"}

    # ---------------- roofline of the dominant kernel (separate pass; events around every launch)
    pk = peaks()
    reps = 5
    prof = sep.profile_kernels(lambda: [sep.separate_batch(mix) for _ in range(reps)])
    top = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else (None, None)
    roofline = None
    tensor_pipe = None
    if top[0]:
        name, rec = top
        bound, work = kernel_work(name, BATCH, T)
        avg_ms = rec["ms"] / rec["launches"]       # one event pair per launch: includes event records, launch gaps, lost PDL overlap
        pair_ms = avg_ms
        method = "one CUDA-event pair around every launch of the kernel inside whole forwards"
        if lk and "error" not in lk:
            key = ("post" if name.startswith("k_post2_tc") else "qkv" if name.startswith("k_qkv2_tc")
                   else "attention" if name.startswith(("k_attention_bf16_tma", "k_attn_tc")) else None)
            if key and not name.endswith("(small)"):
                avg_ms = lk[key] / 1e3
                method = "32 serialised launches of the kernel (no PDL) between one CUDA-event pair, median of 5 (resep_layer_kernel_repeat)"
        frac_sust = None
        if bound == "tensor":
            # the profiling pass times each kernel alone at full clocks: the BURST cuBLAS figure is the honest denominator
            peak, unit = pk["bf16_tflops"], "TFLOP/s"
            achieved = work * reps / rec["launches"] / (avg_ms / 1e3) / 1e12
            frac_sust = achieved / pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        elif bound == "hbm":
            peak, unit = pk["hbm_gbs"], "GB/s"
            achieved = work * reps / rec["launches"] / (avg_ms / 1e3) / 1e9
        else:
            peak = unit = achieved = None
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu capture, averaged per launch
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f).get(name.split("<")[0])
            if tj:
                traffic = tj["dram_bytes_big_launch"]            # the 16 intra-block launches (64,800 rows each)
        except Exception:
            traffic = None
        roofline = {"kernel": name, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                    "peak_source": pk["_src"] + (" (burst bf16: the kernel is timed alone in the profiling pass)" if bound == "tensor" else ""),
                    "frac_of_sustained_peak": frac_sust,
                    "avg_launch_us": 1e3 * avg_ms, "avg_launch_method": method, "event_pair_per_launch_us": 1e3 * pair_ms,
                    "layer_kernels_back_to_back_us": lk,
                    "launches_per_step": rec["launches"] / reps,
                    "share_of_step": rec["ms"] / sum(r["ms"] for r in prof.values()),
                    "kernel_ms_per_step": {k: round(v["ms"] / reps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}
        # the other two layer kernels against their own roofs (back-to-back durations, per launch)
        if lk and "error" not in lk:
            Mrows = shapes(BATCH, T)[3]
            n_ch = shapes(BATCH, T)[2]
            qkv_bytes, qkv_flops = Mrows * (512 + 768), Mrows * 98_304
            att_flops = n_ch * 8 * 4 * 150 * 150 * 16
            att_exp = n_ch * 8 * 150 * 150
            n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
            sm_clock_hz = 1e6 * float(clocks.summary().get("sm_max_mhz") or 1965)     # (the kernels were timed at burst clocks)
            roofline["other_layer_kernels"] = [
                {"kernel": "k_qkv2_tc", "bound": "hbm", "us": lk["qkv"], "achieved": qkv_bytes / (lk["qkv"] * 1e-6) / 1e9, "unit": "GB/s",
                 "peak": pk["hbm_gbs"], "frac": qkv_bytes / (lk["qkv"] * 1e-6) / 1e9 / pk["hbm_gbs"],
                 "tensor_tflops": qkv_flops / (lk["qkv"] * 1e-6) / 1e12,
                 "note": "algorithmic bytes: 512 B in + 768 B out per row; most of them are L2 hits inside the step"},
                {"kernel": "intra attention", "bound": "instruction issue (HMMA + MUFU.EX2 per 16 x 16 block)", "us": lk["attention"],
                 "tensor_tflops": att_flops / (lk["attention"] * 1e-6) / 1e12,
                 "exp2_per_clk_per_sm": att_exp / (lk["attention"] * 1e-6) / (sm_clock_hz * n_sms),
                 "note": "MUFU.EX2 peak 16 per clock per SM (scripts/microbench/attn_probe.cu)"}]
        # all tensor-core work of the intra blocks together: (QKV + attention + out-proj + FFN FLOPs) / their summed time
        tc_flops, tc_ms = 0.0, 0.0
        for k, v in prof.items():
            if k.endswith("(small)") or not (k.startswith("k_qkv") or k.startswith("k_post") or k.startswith("k_attn_tc") or k.startswith("k_attention_bf16_tma")):
                continue
            _, w = kernel_work(k, BATCH, T)
            if k.startswith("k_qkv"):
                w = 16 * shapes(BATCH, T)[3] * 98_304      # (kernel_work reports this HBM-bound kernel in bytes)
            tc_flops += w * reps
            tc_ms += v["ms"]
        if lk and "error" not in lk and tc_ms > 0:      # the layer as the product runs it: 3 PDL launches, 64 layers back to back
            tc_ms = lk["layer_pdl"] / 1e3 * reps * 16
        if tc_ms > 0:
            ach = tc_flops / (tc_ms / 1e3) / 1e12
            tensor_pipe = {"kernels": "k_qkv2_tc + intra attention + k_post2_tc (the 16 intra layers)", "achieved_tflops": ach,
                           "frac_of_burst_peak": ach / pk["bf16_tflops"],
                           "frac_of_sustained_peak": ach / pk.get("bf16_tflops_sustained", pk["bf16_tflops"]),
                           "us_per_layer": 1e3 * tc_ms / reps / 16}

    # ---------------- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if not args.no_cpu_baseline and not args.no_side_blocks and world == 1:
        v, times, threads = time_oracle(sds, items=BATCH, reps=5)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"the whole step (B={BATCH} x {SECONDS} s, coupled), oracle fp32 eager PyTorch, 1 warm-up + median of 5 "
                         f"({sum(times):.1f} s of CPU work)", "host_cpus": os.cpu_count()}
        if args.cpu_other_configs:
            cpu["other_configs"] = time_oracle_other_configs(sds)

    flops_step = None
    from clearconverse_b200.sharding import flops_of
    L, S, chunks, M, mem_seqs = shapes(BATCH, T)
    flops_step = BATCH * flops_of(T, mem_seq_len=chunks)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms.item() / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": bench_config(world),
        "notes": {"weights": "random-init (seed 0); bf16 operands (in-proj / out-proj / output_fc weights as bf16 hi+lo, FFN weights single bf16), fp32 accumulate",
                  "driver": "four forwards in flight on four CUDA streams, each with its own 0.3 GB workspace "
                            "(single_forward: one at a time, 256 MiB L2 flush between steps)"},
        "single_forward": {"value": single_value, "ms_per_step": single_ms.item() / args.steps,
                           "note": "one forward at a time on one stream, per-step CUDA events, L2 flushed between steps"},
        "clocks": clocks.summary(),
        "sustained": sustained,
        "copy_probe": copy_probe,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BATCH * T * 4, "d2h_bytes_per_step": BATCH * T * 2 * 4,
                "api": "SepformerSeparation.separate_stream (pinned host batches in, pinned host results out, copies overlapped, four forwards in flight)",
                "serial_value": e2e_serial_value, "serial_api": "separate_batch(host) + blocking copy per step"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "tensor_pipe": tensor_pipe,
        "cpu_baseline": cpu,
        "meeting": meeting,
        "long_split": long_split,
        "latency": latency,
        "algorithmic_tflops_per_s": flops_step * args.steps * world / total_s / 1e12,
        "wall_s_timed_region": t_wall,
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_meeting(sds, precision, rank, world, dev, dist, barrier, reps: int = 5):
    """BASELINE configs[3]: every overlap segment of a synthetic 1-hour meeting (720 s in 0.5-30 s segments, seed 4),
    per-item semantics, length-bucketed and dealt to the ranks by ``sharding.separate_sharded``.  Every repetition
    starts from the pinned HOST segments and ends with all results in rank 0's host memory: planning, H2D, kernels,
    D2H and the gather are inside the timed region.  The segment list is the same at every N (strong scaling)."""
    import torch
    from clearconverse_b200 import SepformerSeparation, sharding, synth
    lens = synth.meeting_overlap_segments(720.0, seed=4)
    segs = [synth.synth_mixture(n, 5000 + i)[0].pin_memory() for i, n in enumerate(lens)]
    sep_i = SepformerSeparation(sds, device=dev, precision=precision, batch_mode="independent")
    name = f"resep_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
    shared = None
    if rank == 0:
        shared = sharding.SharedResults(lens, name, rank, world)
    barrier()
    if rank != 0:
        shared = sharding.SharedResults(lens, name, rank, world)

    def once(use_shared: bool):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        res, samples = sharding.separate_sharded(segs, sep_i.separate_segments, rank, world, pipeline=sep_i.separate_stream,
                                                 shared=shared if use_shared else None)
        if not use_shared and res is not None:
            res = [r.cpu() if r.is_cuda else r for r in res]       # (world == 1: results to the host like the other form)
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        t = torch.tensor([ev0.elapsed_time(ev1) / 1e3, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), samples, res

    for _ in range(3):
        once(True)
    runs = [once(True) for _ in range(reps)]
    dev_s = statistics.median(r[0] for r in runs)
    wall_s = statistics.median(r[1] for r in runs)
    total_audio = sum(lens) / SAMPLE_RATE
    check = None
    if rank == 0:
        res = runs[-1][3]
        check = bool(all(torch.isfinite(r).all() for r in res[:8]) and res[0].shape == (lens[0], 2))
    once(False)
    pickled = [once(False) for _ in range(3)]
    pick_s = statistics.median(r[1] for r in pickled)
    batches, per_rank = sharding.plan_shards(lens, world)
    out = {"value": total_audio / wall_s, "unit": UNIT, "scaling": "strong", "segments": len(lens), "audio_s": total_audio,
           "ms": 1e3 * wall_s, "ms_device_events": 1e3 * dev_s, "batches": len(batches),
           "batches_per_rank": [len(r) for r in per_rank],
           "gather": "shared page-locked host buffer (sharding.SharedResults): D2H straight to the final offsets, counter barrier",
           "pickle_gather_value": total_audio / pick_s, "pickle_gather_ms": 1e3 * pick_s,
           "timed": "host wall clock, max over ranks, median of %d: plan + H2D + kernels + D2H + gather, from pinned host segments "
                    "to results in rank 0's host memory" % reps,
           "results_ok": check}
    barrier()
    shared.close()
    sep_i.close()
    return out


def run_long_split(sep, rank, world, dev, dist):
    """SURVEY section 8e, optional row: ONE 60 s recording split by chunks across the ranks (the path's only collective
    is an all-gather of the chunk summaries).  Checked bit for bit against the unsplit forward on every rank."""
    import torch
    from clearconverse_b200 import sharding, synth
    mix = synth.synth_mixture(480000, 3)[0].to(dev)
    want = sep.separate_batch(mix[None])[0]
    group = dist.group.WORLD
    got = sharding.separate_long(sep, mix, rank, world, group=group, parts=world)
    flag = torch.tensor([1 if torch.equal(got, want) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    t_split = timed(lambda: sharding.separate_long(sep, mix, rank, world, group=group, parts=world))
    t_one = timed(lambda: sep.separate_batch(mix[None]))
    return {"recording_s": 60, "ranks": world, "bit_identical_on_every_rank": bool(flag.item()), "ms_split": t_split,
            "ms_one_gpu": t_one, "collective": "all_gather of chunk summaries [S,128] fp32 + all_gather of span outputs (NCCL)"}


if __name__ == "__main__":
    main()
