"""ctypes binding of ``libresep_b200.so`` (C ABI declared in ``include/resep_b200.h``).

The library is the product: if it is missing the package fails loudly here -- there is no
PyTorch / CPU fallback anywhere in ``clearconverse_b200``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libresep_b200.so")

ABI_VERSION = 1
PREC_FP32, PREC_TF32, PREC_BF16, PREC_FP16 = 0, 1, 2, 3
BATCH_COUPLED, BATCH_INDEPENDENT = 0, 1
PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "bf16": PREC_BF16, "fp16": PREC_FP16}
BATCH_MODES = {"coupled": BATCH_COUPLED, "independent": BATCH_INDEPENDENT}

E_INVAL, E_SHORT, E_CUDA, E_WORKSPACE, E_POS, E_NODEVICE = -1, -2, -3, -4, -5, -6

_fp = C.POINTER(C.c_float)


class ResepConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("abi_version", "n_filters", "kernel_size", "stride", "segment_size",
                                         "n_heads", "d_ffn", "n_layers", "n_blocks", "n_spks")]


class ResepLayerWeights(C.Structure):
    _fields_ = [(n, _fp) for n in ("norm1_w", "norm1_b", "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b",
                                   "norm2_w", "norm2_b", "ffn1_w", "ffn1_b", "ffn2_w", "ffn2_b")]


class ResepBlockWeights(C.Structure):
    _fields_ = [("layers", ResepLayerWeights * 8), ("final_norm_w", _fp), ("final_norm_b", _fp),
                ("gln_w", _fp), ("gln_b", _fp)]


class ResepWeights(C.Structure):
    _fields_ = [("enc_w", _fp), ("dec_w", _fp), ("prelu_a", _fp), ("fc_w", _fp), ("fc_b", _fp), ("pe", _fp),
                ("pe_rows", C.c_int64), ("seg", ResepBlockWeights * 2), ("mem", ResepBlockWeights * 1)]


class ResepSpanCtl(C.Structure):
    _fields_ = [("phase", C.c_int), ("inner", C.c_int), ("chunk_means", C.c_void_p), ("hc", C.c_void_p)]


class ResepDebugOut(C.Structure):
    _fields_ = [(n, _fp) for n in ("enc", "seg0", "chunk_mean", "mem0", "seg1")]


# every symbol include/resep_b200.h declares: name -> (restype, argtypes)
_i64p = C.POINTER(C.c_int64)
_H = C.c_void_p
SYMBOLS = {
    "resep_create": (C.c_int, [C.POINTER(ResepConfig), C.POINTER(ResepWeights), C.c_int, C.POINTER(_H)]),
    "resep_load_weights": (C.c_int, [_H, C.POINTER(ResepWeights)]),
    "resep_destroy": (C.c_int, [_H]),
    "resep_last_error": (C.c_char_p, [_H]),
    "resep_workspace_bytes": (C.c_int, [_H, C.c_int, _i64p, C.c_int, C.POINTER(C.c_size_t)]),
    "resep_forward": (C.c_int, [_H, C.c_void_p, _i64p, _i64p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                C.c_int, C.c_int, C.c_void_p]),
    "resep_forward_span": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "resep_memory_workspace_bytes": (C.c_int, [_H, C.c_int, C.POINTER(C.c_size_t)]),
    "resep_memory_block": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "resep_resample_fir": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_int, C.c_int, C.c_void_p]),
    "resep_peak_normalize": (C.c_int, [_H, C.c_void_p, _i64p, _i64p, C.c_int, C.c_void_p, C.c_void_p]),
    "resep_forward_debug": (C.c_int, [_H, C.c_void_p, _i64p, _i64p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_int, C.c_int, C.c_void_p, C.POINTER(ResepDebugOut)]),
    "resep_launch_count": (C.c_int64, [_H]),
    "resep_profile": (C.c_int, [_H, C.c_int]),
    "resep_profile_report": (C.c_int, [_H, C.c_char_p, C.c_size_t]),
    "resep_encoder_fwd": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "resep_layer_fwd": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                  C.c_int, C.c_void_p]),
    "resep_layer_kernel_repeat": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                            C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "resep_linear_fwd": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree CUDA library; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built "
            "(run `python -m clearconverse_b200.build`).  clearconverse_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError here == ABI drift: fail loudly
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class ResepError(RuntimeError):
    """A non-zero status from the C ABI.  It is a RuntimeError so that the reference's
    ``except Exception`` at /root/reference/back/api.py:1107 handles it like upstream's errors."""

    def __init__(self, code: int, message: str):
        super().__init__(f"resep_b200 error {code}: {message}")
        self.code = code


def check(lib, handle, rc: int):
    if rc != 0:
        msg = lib.resep_last_error(handle)
        raise ResepError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")
