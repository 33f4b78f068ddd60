"""Drop-in for ``speechbrain.inference.separation.SepformerSeparation`` on the one path
ClearConverse uses it for (/root/reference/back/api.py):

    self.separator = SepformerSeparation.from_hparams(source=..., savedir=..., run_opts={"device": dev})   # :713-717
    self.separator.load_state_dict(state_dict, strict=False)                                              # :745
    separated = self.separator.separate_batch(subsegment)        # [1,T] -> [1,T,2]                       # :1077

Same names, argument meaning and error behaviour (Python exceptions; the caller's
``except Exception`` at api.py:1107 turns them into "[Processing error]" rows).  Behind it a
PyTorch custom op (CUDA dispatch key only) hands raw device pointers and the current CUDA
stream to the C ABI of ``libresep_b200.so``.  There is no CPU path: a CPU-only host or a
missing library raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import weakref
from collections import namedtuple
from types import SimpleNamespace

import torch

from . import _lib, weights as _weights

NUM_SPKS = 2
SAMPLE_RATE = 8000
KERNEL_SIZE = 16

_IncompatibleKeys = namedtuple("_IncompatibleKeys", ["missing_keys", "unexpected_keys"])

# ---------------------------------------------------------------------------------------------
# engine registry: the custom op takes an integer engine id (ops cannot carry Python objects).  Weak values: the
# registry must not keep an engine (its device weights, workspaces, static I/O buffers, CUDA graphs) alive once the
# separator that owns it has been dropped -- upstream's object frees its memory when it goes out of scope.
_ENGINES: "weakref.WeakValueDictionary[int, _Engine]" = weakref.WeakValueDictionary()
_ENGINES_LOCK = threading.Lock()
_NEXT_ID = [1]
# The blocking calls (separate_batch / separate_segments) own this workspace lane and static-I/O slot; the lanes of
# the pipelined driver are 0 and 1, its slots 0 .. depth-1.  Keeping the namespaces apart means a separate_batch
# issued on the caller's stream while a separate_stream generator is half consumed (or was abandoned) never touches
# a buffer or workspace that a lane stream may still be using.
SYNC_LANE = "sync"
# compute lanes of separate_stream (CUDA streams with their own workspace and graph set); slot s of the pipeline runs on lane s % N_LANES
N_LANES = max(1, min(4, int(os.environ.get("RESEP_LANES", "4"))))


def _destroy_handle(lib, handle):
    if handle is not None and handle.value:
        lib.resep_destroy(handle)
        handle.value = None


class _Engine:
    """Owns one ResepHandle (one CUDA device) and a grow-only workspace tensor."""

    def __init__(self, sds: dict, device: torch.device, pe_rows: int):
        if device.type != "cuda":
            raise RuntimeError(
                f"clearconverse_b200 runs on CUDA devices only (got '{device}'); there is no CPU fallback")
        if not torch.cuda.is_available():
            raise RuntimeError("clearconverse_b200: no CUDA device is available; there is no CPU fallback")
        self.lib = _lib.load_library()
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        self.packed = _weights.PackedWeights(sds, pe_rows)
        self.handle = C.c_void_p()
        cfg = _weights.default_config()
        rc = self.lib.resep_create(C.byref(cfg), C.byref(self.packed.struct), self.device.index, C.byref(self.handle))
        _lib.check(self.lib, None, rc)
        self.workspaces: dict = {}        # lane -> tensor (a lane = one CUDA stream's worth of in-flight forwards)
        self._static_io: dict = {}        # (offs, lens, slot) -> (mix_buf, est_buf): fixed addresses let the C side replay a CUDA graph
        with _ENGINES_LOCK:
            self.id = _NEXT_ID[0]
            _NEXT_ID[0] += 1
            _ENGINES[self.id] = self
        # belt and braces next to __del__: runs at interpreter exit too, and never resurrects the engine
        self._finalizer = weakref.finalize(self, _destroy_handle, self.lib, self.handle)

    def reload(self, sds: dict, pe_rows: int):
        packed = _weights.PackedWeights(sds, pe_rows)
        _lib.check(self.lib, self.handle, self.lib.resep_load_weights(self.handle, C.byref(packed.struct)))
        self.packed = packed

    def launch_count(self) -> int:
        return int(self.lib.resep_launch_count(self.handle))

    def _workspace_for(self, lens: "C.Array", B: int, precision: int, lane: int = 0) -> torch.Tensor:
        need = C.c_size_t()
        _lib.check(self.lib, self.handle, self.lib.resep_workspace_bytes(self.handle, B, lens, precision, C.byref(need)))
        ws = self.workspaces.get(lane)
        if ws is None or ws.numel() < need.value:
            self.workspaces.pop(lane, None)          # release before growing
            ws = torch.empty(int(need.value * 1.0) + 1024, dtype=torch.uint8, device=self.device)
            self.workspaces[lane] = ws
        return ws

    def static_io(self, offs, lens, total: int, slot: int = 0):
        """Persistent (mix, est) device buffers for one batch shape.  The C ABI keys its CUDA graphs on buffer
        addresses, so feeding the same buffers makes every call after the second a single graph launch."""
        key = (tuple(offs), tuple(lens), total, slot)
        io = self._static_io.get(key)
        if io is None:
            if len(self._static_io) >= 24:
                old_key = next(iter(self._static_io))
                old = self._static_io.pop(old_key)
                if old_key[3] == SYNC_LANE:               # used on the caller's stream only
                    for t in old:
                        t.record_stream(torch.cuda.current_stream(self.device))
                else:                                     # a pipeline slot: its lane stream may still be using it
                    torch.cuda.synchronize(self.device)
            with torch.cuda.device(self.device):
                io = (torch.empty(total, dtype=torch.float32, device=self.device),
                      torch.zeros(2 * total, dtype=torch.float32, device=self.device))
            self._static_io[key] = io
        return io

    def forward(self, mix_flat: torch.Tensor, offs: list[int], lens: list[int], precision: int, batch_mode: int,
                debug: dict | None = None, out: torch.Tensor | None = None, lane=SYNC_LANE) -> torch.Tensor:
        """mix_flat: 1-D fp32 CUDA tensor holding every item; returns est_flat [2 * mix_flat.numel()]
        (``out`` if given: it must be zero-filled where items leave gaps).  Forwards that may be in flight at the same
        time (different CUDA streams) must use different ``lane``s: a lane owns a workspace."""
        B = len(lens)
        c_off = (C.c_int64 * B)(*offs)
        c_len = (C.c_int64 * B)(*lens)
        with torch.cuda.device(self.device):
            ws = self._workspace_for(c_len, B, precision, lane)
            if out is not None:
                est = out
            else:
                est = torch.zeros(2 * mix_flat.numel(), dtype=torch.float32, device=self.device) \
                    if _needs_zero_fill(offs, lens, mix_flat.numel()) else \
                    torch.empty(2 * mix_flat.numel(), dtype=torch.float32, device=self.device)
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            if debug is None:
                rc = self.lib.resep_forward(self.handle, mix_flat.data_ptr(), c_off, c_len, B, est.data_ptr(),
                                            ws.data_ptr(), ws.numel(), precision, batch_mode, stream)
            else:
                chunks = sum(((t - 16) // 8 + 1) // 150 + 1 for t in lens)
                for name, rows in (("enc", chunks * 150), ("seg0", chunks * 150), ("chunk_mean", chunks),
                                   ("mem0", chunks), ("seg1", chunks * 150)):
                    debug[name] = torch.empty(rows, 128, dtype=torch.float32, device=self.device)
                dbg = _lib.ResepDebugOut(*[C.cast(debug[n].data_ptr(), C.POINTER(C.c_float))
                                           for n in ("enc", "seg0", "chunk_mean", "mem0", "seg1")])
                rc = self.lib.resep_forward_debug(self.handle, mix_flat.data_ptr(), c_off, c_len, B, est.data_ptr(),
                                                  ws.data_ptr(), ws.numel(), precision, batch_mode, stream,
                                                  C.byref(dbg))
            _lib.check(self.lib, self.handle, rc)
        return est

    # ---- one recording split by chunks (sharding.separate_long): phases of resep_forward_span + the memory block
    def _span_call(self, phase: int, mix_ptr: int, span_len: int, est_ptr: int, inner: bool, precision: str, lane: int,
                   means_ptr: int = 0, hc_ptr: int = 0):
        lens = (C.c_int64 * 1)(span_len)
        prec = _lib.PRECISIONS[precision]
        with torch.cuda.device(self.device):
            ws = self._workspace_for(lens, 1, prec, ("span", lane))
            ctl = _lib.ResepSpanCtl(phase, 1 if inner else 0, means_ptr, hc_ptr)
            rc = self.lib.resep_forward_span(self.handle, mix_ptr, span_len, est_ptr, ws.data_ptr(), ws.numel(), prec,
                                             C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream), C.byref(ctl))
        _lib.check(self.lib, self.handle, rc)

    def span_phase1(self, mix_span: torch.Tensor, inner: bool, precision: str, lane: int = 0) -> torch.Tensor:
        """encoder + seg_model[0] on one span (device tensor [span_len]) -> chunk summaries [n_chunks, 128]."""
        n = int(mix_span.numel())
        L = (n - KERNEL_SIZE) // 8 + 1
        n_chunks = L // 150 if (inner and L % 150 == 0) else L // 150 + 1
        means = torch.empty(n_chunks, 128, dtype=torch.float32, device=self.device)
        dummy = torch.empty(2, dtype=torch.float32, device=self.device)
        self._span_call(1, mix_span.data_ptr(), n, dummy.data_ptr(), inner, precision, lane, means_ptr=means.data_ptr())
        return means

    def span_phase2(self, span_len: int, hc: torch.Tensor, inner: bool, precision: str, lane: int = 0) -> torch.Tensor:
        """seg_model[1](out + hc) + mask + decoder of the span whose phase 1 ran in the same lane -> est [span_len, 2]."""
        est = torch.zeros(span_len, NUM_SPKS, dtype=torch.float32, device=self.device)
        dummy = torch.empty(2, dtype=torch.float32, device=self.device)
        self._span_call(2, dummy.data_ptr(), span_len, est.data_ptr(), inner, precision, lane, hc_ptr=hc.data_ptr())
        return est

    def memory_block(self, chunk_means: torch.Tensor, precision: str) -> torch.Tensor:
        """mem_model[0] over one sequence of chunk summaries [S, 128] -> [S, 128]."""
        S = int(chunk_means.shape[0])
        need = C.c_size_t()
        _lib.check(self.lib, self.handle, self.lib.resep_memory_workspace_bytes(self.handle, S, C.byref(need)))
        hc = torch.empty_like(chunk_means)
        with torch.cuda.device(self.device):
            ws = self.workspaces.get("mem")
            if ws is None or ws.numel() < need.value:
                ws = torch.empty(need.value + 1024, dtype=torch.uint8, device=self.device)
                self.workspaces["mem"] = ws
            rc = self.lib.resep_memory_block(self.handle, chunk_means.contiguous().data_ptr(), hc.data_ptr(), S, ws.data_ptr(), ws.numel(),
                                             _lib.PRECISIONS[precision], C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _lib.check(self.lib, self.handle, rc)
        return hc

    def resample(self, x: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
        """x [rows, n, channels] fp32 on the device -> [rows, ceil(n * new / orig), channels]; torchaudio.functional.resample's
        arithmetic (its zero padding included) as one FIR pass on the device."""
        key = (int(orig_freq), int(new_freq))
        cache = self.__dict__.setdefault("_taps", {})
        if key not in cache:
            taps, width, orig, new = resample_taps(*key)
            cache[key] = (taps.to(self.device), width, orig, new)
        taps, width, orig, new = cache[key]
        rows, n, ch = x.shape
        n_out = -(-n * new // orig)
        y = torch.empty(rows, n_out, ch, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.resep_resample_fir(self.handle, x.data_ptr(), rows, n, y.data_ptr(), n_out, ch, orig, new, taps.data_ptr(),
                                             taps.shape[1], width, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _lib.check(self.lib, self.handle, rc)
        return y

    def peak_normalize(self, est_flat: torch.Tensor, offs: list[int], lens: list[int]) -> torch.Tensor:
        """In place: every (item, speaker) source of est_flat divided by its max |.| + 1e-8 (api.py:1082).
        Returns the maxima, [B, n_spk] on the device."""
        B = len(lens)
        peaks = torch.empty(B, NUM_SPKS, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.resep_peak_normalize(self.handle, est_flat.data_ptr(), (C.c_int64 * B)(*offs), (C.c_int64 * B)(*lens),
                                               B, peaks.data_ptr(), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _lib.check(self.lib, self.handle, rc)
        return peaks

    def close(self):
        fin = getattr(self, "_finalizer", None)
        if fin is not None and fin.alive:
            fin()                                     # resep_destroy, exactly once
        self.handle = C.c_void_p()
        self.workspaces = {}
        self._static_io = {}
        with _ENGINES_LOCK:
            _ENGINES.pop(getattr(self, "id", -1), None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def resample_taps(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """Windowed-sinc polyphase taps exactly as torchaudio.functional.resample (method "sinc_interp_hann") defines them:
    returns (taps [new, 2 * width + orig] float32 on the CPU, width, orig, new) with orig / new reduced by their gcd."""
    import math
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1)[:, None, None] / new + idx      # (int64 / int -> float32, as torchaudio computes it)
    t *= base
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t)
    kernels *= window * (base / orig)
    return kernels.to(torch.float32).reshape(new, -1).contiguous(), width, orig, new


def _load_audio(path: str):
    """(waveform [channels, T] float32 in [-1, 1], sample rate).  ``torchaudio.load`` as upstream; torchaudio >= 2.9
    needs the optional torchcodec package for that, so PCM / float WAV files fall back to scipy's reader."""
    try:
        import torchaudio
        return torchaudio.load(path)
    except (ImportError, RuntimeError, OSError):
        pass
    import numpy as np
    from scipy.io import wavfile
    fs, data = wavfile.read(path)
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    x = torch.from_numpy(np.ascontiguousarray(x))
    return (x[None] if x.dim() == 1 else x.T.contiguous()), int(fs)


def _needs_zero_fill(offs, lens, total) -> bool:
    """True when the items do not tile the flat buffer exactly (gaps would stay uninitialised)."""
    pos = 0
    for o, n in sorted(zip(offs, lens)):
        if o != pos:
            return True
        pos = o + n
    return pos != total


# ---------------------------------------------------------------------------------------------
# the thin PyTorch custom op (CUDA dispatch key only -- a CPU tensor finds no kernel and raises)
@torch.library.custom_op("clearconverse_b200::resep_separate", mutates_args=(), device_types="cuda")
def _resep_separate(mix_flat: torch.Tensor, offs: list[int], lens: list[int], engine: int, precision: int,
                    batch_mode: int) -> torch.Tensor:
    eng = _ENGINES.get(engine)
    if eng is None:
        raise RuntimeError("clearconverse_b200: separator engine was destroyed")
    mix_buf, est_buf = eng.static_io(offs, lens, mix_flat.numel(), SYNC_LANE)
    mix_buf.copy_(mix_flat)
    eng.forward(mix_buf, offs, lens, precision, batch_mode, out=est_buf, lane=SYNC_LANE)
    return est_buf.clone()                      # a new tensor owned by the caller, as upstream returns


@torch.library.custom_op("clearconverse_b200::resep_separate_static", mutates_args=(), device_types="cuda")
def _resep_separate_static(host_or_dev_mix: torch.Tensor, offs: list[int], lens: list[int], engine: int, precision: int,
                           batch_mode: int, slot: int) -> torch.Tensor:
    """Like resep_separate, but returns the engine's persistent output buffer of `slot` itself (no clone): the
    pipelined driver copies it to the host and only reuses the slot after that copy has completed."""
    eng = _ENGINES.get(engine)
    if eng is None:
        raise RuntimeError("clearconverse_b200: separator engine was destroyed")
    _, est_buf = eng.static_io(offs, lens, host_or_dev_mix.numel(), slot)
    eng.forward(host_or_dev_mix, offs, lens, precision, batch_mode, out=est_buf, lane=slot % N_LANES)
    return est_buf


@_resep_separate_static.register_fake
def _(host_or_dev_mix, offs, lens, engine, precision, batch_mode, slot):
    return host_or_dev_mix.new_empty(2 * host_or_dev_mix.numel())


@_resep_separate.register_fake
def _(mix_flat, offs, lens, engine, precision, batch_mode):
    return mix_flat.new_empty(2 * mix_flat.numel())


# ---------------------------------------------------------------------------------------------
class _Component:
    """Stand-in for ``separator.mods.<name>``: holds that component's state dict."""

    def __init__(self, owner, name):
        self._owner, self._name = owner, name

    def state_dict(self):
        return dict(self._owner._sds[self._name])


class SepformerSeparation:
    """B200-native RE-SepFormer separator with upstream's inference surface.

    precision: "fp16" (default: the fused tcgen05 kernels with IEEE fp16 operands and fp16 hi + lo weights -- tf32-class
    accuracy, max-abs <= 1e-3 against the fp32 reference, at 4x the speed of the "tf32" mode), "bf16" (the same kernels
    with bf16 operands: the throughput mode, SI-SNR delta <= 0.05 dB), "tf32" (kind::tf32 GEMMs, unfused) or "fp32"
    (FMA kernels, no tensor cores: the reference-grade path);
    batch_mode: "coupled" = upstream's literal batched semantics, "independent" = per item
    (== looping B=1, what the product does).  For B == 1 the two coincide.
    """

    def __init__(self, state_dicts: dict, device="cuda", precision: str | None = None,
                 batch_mode: str = "coupled", pe_rows: int = _weights.DEFAULT_PE_ROWS):
        precision = precision or os.environ.get("RESEP_PRECISION", "fp16")
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {list(_lib.PRECISIONS)}")
        if batch_mode not in _lib.BATCH_MODES:
            raise ValueError(f"batch_mode must be one of {list(_lib.BATCH_MODES)}")
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None and torch.cuda.is_available():
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.precision, self.batch_mode, self._pe_rows = precision, batch_mode, pe_rows
        self._sds = {c: {k: v.detach().to("cpu", torch.float32).clone() for k, v in sd.items()
                         if not k.endswith("pos_enc.pe")} for c, sd in state_dicts.items()}
        self._engine = _Engine(self._sds, self.device, pe_rows)
        self.hparams = SimpleNamespace(num_spks=NUM_SPKS, sample_rate=SAMPLE_RATE)
        self.mods = SimpleNamespace(encoder=_Component(self, "encoder"), masknet=_Component(self, "masknet"),
                                    decoder=_Component(self, "decoder"))

    # ----------------------------------------------------------------------------- construction
    @classmethod
    def from_hparams(cls, source: str, savedir: str | None = None, run_opts: dict | None = None,
                     hparams_file: str = "hyperparams.yaml", random_init_seed: int | None = None, **kwargs):
        """Mirror of ``Pretrained.from_hparams`` for api.py:713-717.  Checkpoints are looked up in
        ``savedir`` then in ``source`` (when it is a local directory).  There is no network on
        the target hosts, so a hub id with no local copy raises unless ``random_init_seed`` (or
        the env var RESEP_RANDOM_INIT_SEED) explicitly asks for random weights."""
        device = (run_opts or {}).get("device", "cuda" if torch.cuda.is_available() else "cpu")
        sds = None
        for d in (savedir, source):
            if d and os.path.isdir(d):
                sds = _weights.load_checkpoint_dir(d)
                if sds is not None:
                    break
        if sds is None:
            seed = random_init_seed if random_init_seed is not None else os.environ.get("RESEP_RANDOM_INIT_SEED")
            if seed is None:
                raise FileNotFoundError(
                    f"no encoder.ckpt/masknet.ckpt/decoder.ckpt under savedir={savedir!r} or source={source!r} "
                    "and no network to fetch them; pass random_init_seed=... for random weights")
            sds = _weights.random_init_state_dicts(int(seed))
        return cls(sds, device=device, precision=kwargs.get("precision"),
                   batch_mode=kwargs.get("batch_mode", "coupled"))

    # ------------------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict: dict, strict: bool = True):
        """``nn.Module.load_state_dict`` semantics over keys ``mods.<component>.<param>``.

        api.py:738-745 passes ``{'masknet': sd, 'encoder': sd, 'decoder': sd}`` with
        ``strict=False``; for upstream's nn.Module those three keys are simply *unexpected* and
        silently ignored, so this must not raise and changes nothing.  Use
        ``load_component_state_dicts`` to really apply fine-tuned checkpoints."""
        own = {f"mods.{c}.{k}": (c, k) for c, sd in self._sds.items() for k in sd}
        unexpected = [k for k in state_dict if k not in own and not k.endswith("pos_enc.pe")]
        missing = [k for k in own if k not in state_dict]
        if strict and (unexpected or missing):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {len(missing)} key(s), "
                               f"unexpected key(s) {unexpected[:5]}")
        hits = {k: v for k, v in state_dict.items() if k in own}
        if hits:
            new = {c: dict(sd) for c, sd in self._sds.items()}
            for k, v in hits.items():
                c, name = own[k]
                if tuple(v.shape) != tuple(new[c][name].shape):
                    raise RuntimeError(f"size mismatch for {k}")
                new[c][name] = v.detach().to("cpu", torch.float32).clone()
            self._apply(new)
        return _IncompatibleKeys(missing, unexpected)

    def load_component_state_dicts(self, state_dicts: dict):
        """Apply ``{'encoder': sd, 'masknet': sd, 'decoder': sd}`` (any subset) for real."""
        new = {c: dict(sd) for c, sd in self._sds.items()}
        for c, sd in state_dicts.items():
            if c not in new:
                raise KeyError(c)
            for k, v in sd.items():
                if k.endswith("pos_enc.pe"):
                    continue
                new[c][k] = v.detach().to("cpu", torch.float32).clone()
        self._apply(new)

    def _apply(self, new):
        _weights.validate_state_dicts(new)
        self._engine.reload(new, self._pe_rows)
        self._sds = new

    def state_dict(self):
        return {f"mods.{c}.{k}": v for c, sd in self._sds.items() for k, v in sd.items()}

    # ------------------------------------------------------------------------------ inference
    def _check_mix(self, mix):
        if not isinstance(mix, torch.Tensor):
            raise TypeError("mix must be a torch.Tensor")
        if mix.dim() != 2:
            raise RuntimeError(f"Expected a [batch, time] mixture, got shape {tuple(mix.shape)}")
        if mix.dtype != torch.float32:
            raise RuntimeError(f"expected a float32 mixture (the model weights are float32), got {mix.dtype}")
        if mix.size(0) == 0:
            raise RuntimeError("empty batch")
        if mix.size(1) < KERNEL_SIZE:
            raise RuntimeError(f"Calculated padded input size per channel: ({mix.size(1)}). Kernel size: (16). "
                               "Kernel size can't be greater than actual input size")

    @torch.no_grad()
    def separate_batch(self, mix: torch.Tensor, peak_normalize: bool = False, sample_rate: int | None = None) -> torch.Tensor:
        """mix [B,T] float32 (any device) -> est_sources [B,T,n_spk] float32 on ``self.device``.
        ``peak_normalize=True`` additionally applies the caller's `source / (source.abs().max() + 1e-8)`
        (api.py:1082) to every (item, speaker) source on the device.
        ``sample_rate`` (default: the model's 8000, i.e. no resampling -- what api.py does with its 16 kHz audio):
        the rate of ``mix``; if it differs, mix is resampled to 8 kHz, separated, and the sources are resampled back,
        all on the device, so the result is still [B,T,n_spk] at the caller's rate."""
        self._check_mix(mix)
        if sample_rate is not None and int(sample_rate) != SAMPLE_RATE:
            B, T = mix.shape
            x = self._engine.resample(mix.to(self.device).contiguous().view(B, T, 1), int(sample_rate), SAMPLE_RATE)
            est = self.separate_batch(x.view(B, -1))
            est = self._engine.resample(est, SAMPLE_RATE, int(sample_rate))[:, :T].contiguous()
            if est.size(1) < T:
                est = torch.nn.functional.pad(est, (0, 0, 0, T - est.size(1)))
            if peak_normalize:
                self._engine.peak_normalize(est.view(-1), [b * T for b in range(B)], [T] * B)
            return est
        B, T = mix.shape
        mix = mix.to(self.device).contiguous()
        offs, lens = [b * T for b in range(B)], [T] * B
        est = torch.ops.clearconverse_b200.resep_separate(
            mix.view(-1), offs, lens, self._engine.id,
            _lib.PRECISIONS[self.precision], _lib.BATCH_MODES[self.batch_mode])
        if peak_normalize:
            self._engine.peak_normalize(est, offs, lens)
        return est.view(B, T, NUM_SPKS)

    forward = separate_batch
    __call__ = separate_batch

    @torch.no_grad()
    def separate_segments(self, segments: list[torch.Tensor], peak_normalize: bool = False) -> list[torch.Tensor]:
        """Ragged batch: 1-D (or [1,T]) float32 segments of different lengths in ONE launch
        sequence, each separated independently (== one ``separate_batch`` call per segment, the
        way api.py:1073-1077 loops).  Returns a list of [T_i, n_spk] tensors on ``self.device``."""
        flat, offs, lens, pos = [], [], [], 0
        for s in segments:
            s = s.reshape(-1)
            if s.dtype != torch.float32:
                raise RuntimeError(f"expected float32 segments, got {s.dtype}")
            if s.numel() < KERNEL_SIZE:
                raise RuntimeError("Kernel size can't be greater than actual input size")
            flat.append(s.to(self.device, non_blocking=True))
            offs.append(pos)
            lens.append(s.numel())
            pos += s.numel()
        if not flat:
            return []
        est = torch.ops.clearconverse_b200.resep_separate(
            torch.cat(flat), offs, lens, self._engine.id, _lib.PRECISIONS[self.precision], _lib.BATCH_INDEPENDENT)
        if peak_normalize:
            self._engine.peak_normalize(est, offs, lens)
        return [est[2 * o:2 * (o + n)].view(n, NUM_SPKS) for o, n in zip(offs, lens)]

    @torch.no_grad()
    def separate_regions(self, audio: torch.Tensor, spans: list[tuple[int, int]], peak_normalize: bool = False) -> list[torch.Tensor]:
        """The batched overlap driver: ``audio`` is one file's waveform ([T] or [1,T]),
        ``spans`` the (start, end) sample ranges of its overlap regions in the caller's order -- what api.py:1073-1077
        slices one at a time.  Identical spans (two speakers' segments over the same overlap, api.py:1387-1394) are
        separated once.  Returns one [T_i, n_spk] tensor per span (duplicates share storage)."""
        from .sharding import dedupe_spans
        audio = audio.reshape(-1)
        unique, inverse = dedupe_spans(spans)
        outs = self.separate_segments([audio[a:b] for a, b in unique], peak_normalize=peak_normalize)
        return [outs[i] for i in inverse]

    @torch.no_grad()
    def separate_stream(self, batches, out_buffers=None, depth: int = 2, device_out: bool = False):
        """Pipelined driver for a sequence of HOST batches (the batched overlap driver of SURVEY.md section 8f-1):
        yields, in order, a pinned HOST tensor [B,T,n_spk] per input batch.  The host->device copy of batch i+1 and
        the device->host copy of batch i-1 run on their own streams under the kernels of batch i, so the copies cost
        no throughput.  ``batches``: iterable of [B,T] float32 CPU tensors (pinned for truly asynchronous copies);
        ``out_buffers``: optional list of >= ``depth`` pinned [B,T,n_spk] tensors to reuse (a yielded tensor is
        overwritten ``depth`` batches later).  Batches that already live on the device are taken as they are (no host
        copy); ``device_out=True`` yields device tensors (new ones, owned by the caller) instead of host tensors.  A batch
        may also be a LIST of 1-D segments of different lengths (device or pinned host): it is separated with per-item
        semantics like ``separate_segments`` and yields a list of [T_i, n_spk] device tensors."""
        dev = self.device
        # min(depth, N_LANES) compute streams, batches dealt to them in turn (each lane has its own workspace and CUDA
        # graphs): a forward spends ~12 % of its time in the memory transformer, whose latency-bound launches use few
        # CTAs, and every persistent layer kernel ends in a partly filled last round; the neighbouring batches' kernels
        # fill those SMs (scripts/gpu_timeline.py, scripts/gpu_dual_stream.py; config 2, scripts/gpu_lanes.py: 1 lane
        # 40.4 k audio-s/s, 2 lanes 43.1-43.5 k, 3 lanes 43.7-44.1 k, 4 lanes 44.3 k).
        if getattr(self, "_pipe_streams", None) is None:   # created once: the caching allocator keeps per-stream pools,
            self._pipe_streams = [torch.cuda.Stream(dev) for _ in range(N_LANES + 2)]   # fresh streams mean fresh cudaMallocs
        lanes = self._pipe_streams[:max(1, min(depth, N_LANES))]
        h2d, d2h = self._pipe_streams[N_LANES], self._pipe_streams[N_LANES + 1]
        caller = torch.cuda.current_stream(dev)
        for st in self._pipe_streams:
            st.wait_stream(caller)
        inflight = []                                     # (event, host_out)

        def drain_one():
            ev, out = inflight.pop(0)
            ev.synchronize()
            return out

        def hand_over(t):
            # device results are allocated from the lane stream's pool but consumed on the caller's stream: tell the
            # caching allocator, or the block could be handed out again on the lane under a pending reader
            t.record_stream(caller)
            return t

        try:
            for i, mix in enumerate(batches):
                slot = i % depth
                compute = lanes[slot % N_LANES % len(lanes)]     # the lane (workspace) of a slot is slot % N_LANES
                if isinstance(mix, (list, tuple)):
                    # a ragged batch of 1-D segments (per-item semantics, as separate_segments): results stay on the device
                    segs = [s_.reshape(-1) for s_ in mix]
                    for s_ in segs:
                        if s_.dtype != torch.float32 or s_.numel() < KERNEL_SIZE:
                            raise RuntimeError("expected float32 segments of at least 16 samples")
                    lens = [int(s_.numel()) for s_ in segs]
                    offs = [0] * len(lens)
                    for j in range(1, len(lens)):
                        offs[j] = offs[j - 1] + lens[j - 1]
                    total = offs[-1] + lens[-1]
                    mix_buf, _ = self._engine.static_io(offs, lens, total, slot)
                    with torch.cuda.stream(compute):
                        for o_, n_, s_ in zip(offs, lens, segs):
                            mix_buf[o_:o_ + n_].copy_(s_, non_blocking=True)
                        est = torch.ops.clearconverse_b200.resep_separate_static(
                            mix_buf, offs, lens, self._engine.id, _lib.PRECISIONS[self.precision], _lib.BATCH_INDEPENDENT, slot)
                        res = hand_over(est.clone())
                        fin = torch.cuda.Event(); fin.record(compute)
                    inflight.append((fin, [res[2 * o_:2 * (o_ + n_)].view(n_, NUM_SPKS) for o_, n_ in zip(offs, lens)]))
                    if len(inflight) >= depth:
                        yield drain_one()
                    continue
                self._check_mix(mix)
                B, T = mix.shape
                offs, lens = [b * T for b in range(B)], [T] * B
                # slot i % depth was last used by batch i - depth, whose result has been drained (synchronised) already
                mix_buf, _ = self._engine.static_io(offs, lens, B * T, slot)
                if mix.is_cuda:
                    with torch.cuda.stream(compute):
                        mix_buf.copy_(mix.reshape(-1), non_blocking=True)
                else:
                    with torch.cuda.stream(h2d):
                        mix_buf.copy_(mix.reshape(-1), non_blocking=True)
                        up = torch.cuda.Event(); up.record(h2d)
                    compute.wait_event(up)
                with torch.cuda.stream(compute):
                    est = torch.ops.clearconverse_b200.resep_separate_static(
                        mix_buf, offs, lens, self._engine.id, _lib.PRECISIONS[self.precision],
                        _lib.BATCH_MODES[self.batch_mode], slot).view(B, T, NUM_SPKS)
                if device_out:
                    with torch.cuda.stream(compute):
                        res = hand_over(est.clone())
                        fin = torch.cuda.Event(); fin.record(compute)
                    inflight.append((fin, res))
                    if len(inflight) >= depth:
                        yield drain_one()
                    continue
                done = torch.cuda.Event(); done.record(compute)
                if out_buffers is not None:
                    host = out_buffers[i % len(out_buffers)]
                else:
                    host = torch.empty(B, T, NUM_SPKS, dtype=torch.float32).pin_memory()
                with torch.cuda.stream(d2h):
                    d2h.wait_event(done)
                    host.copy_(est, non_blocking=True)
                    down = torch.cuda.Event(); down.record(d2h)
                inflight.append((down, host))
                if len(inflight) >= depth:
                    yield drain_one()
            while inflight:
                yield drain_one()
        finally:
            # also on an abandoned generator or an exception: whatever the caller issues next on its stream is ordered
            # after everything the lanes still have in flight (a device-side wait; the host does not block)
            for st in self._pipe_streams:
                caller.wait_stream(st)

    def separate_batch_debug(self, mix: torch.Tensor) -> tuple[torch.Tensor, dict]:
        """separate_batch + intermediates (encoder / block outputs) for per-kernel parity tests."""
        self._check_mix(mix)
        B, T = mix.shape
        mix = mix.to(self.device).contiguous()
        dbg: dict = {}
        est = self._engine.forward(mix.view(-1), [b * T for b in range(B)], [T] * B,
                                   _lib.PRECISIONS[self.precision], _lib.BATCH_MODES[self.batch_mode], dbg)
        return est.view(B, T, NUM_SPKS), dbg

    def separate_file(self, path: str, savedir: str | None = None) -> torch.Tensor:
        """Upstream's convenience wrapper (speechbrain/inference/separation.py ``separate_file``): load, mono-mix and
        resample to the model's 8 kHz when the file's rate differs, separate, and divide by the per-source peak.
        Returns [1, T, n_spk] on ``self.device``.  The resampler is the device FIR pass (torchaudio's taps)."""
        batch, fs = _load_audio(path)
        batch = batch.to(self.device)
        if fs != SAMPLE_RATE:
            batch = batch.mean(dim=0, keepdim=True)
            batch = self._engine.resample(batch.float().contiguous()[:, :, None], fs, SAMPLE_RATE)[:, :, 0]
        est = self.separate_batch(batch.float().contiguous())
        return est / est.abs().max(dim=1, keepdim=True)[0]

    def launch_count(self) -> int:
        return self._engine.launch_count()

    def profile_kernels(self, fn) -> dict:
        """Run ``fn()`` with every kernel launch bracketed by CUDA events on its stream and
        return {kernel name: {"ms": total, "launches": n}} (bench.py's roofline source)."""
        import json
        eng = self._engine
        _lib.check(eng.lib, eng.handle, eng.lib.resep_profile(eng.handle, 1))
        try:
            fn()
        finally:
            buf = C.create_string_buffer(1 << 16)
            _lib.check(eng.lib, eng.handle, eng.lib.resep_profile_report(eng.handle, buf, len(buf)))
        return json.loads(buf.value.decode())

    def close(self):
        eng = getattr(self, "_engine", None)
        if eng is not None:
            eng.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        # upstream's object releases its GPU memory when dropped; so does this one (device weights, workspaces,
        # static I/O buffers, plans and CUDA graphs all hang off the engine)
        try:
            self.close()
        except Exception:
            pass
