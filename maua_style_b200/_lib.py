"""ctypes binding of libmaua_b200.so (the C ABI declared in include/maua_b200.h).

There is no fallback: if the shared library is missing it is built in-tree with nvcc; if that is impossible, or
a compute entry point is called without an sm_100 GPU, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libmaua_b200.so"
HEADER_PATH = _HERE.parent / "include" / "maua_b200.h"

MAUA_IMPL_TC = 0
MAUA_IMPL_REF = 1
MAUA_IMPL_TC_1CTA = 2
MAUA_IMPL_TC_2CTA = 3
MAUA_IMPL_FP32 = 4  # exact arithmetic (csrc/conv_fp32.cu): parity checks and MAUA_PRECISION=fp32
MAUA_MAX_LAYERS = 32
MAUA_MAX_TAPS = 16
MODE_NONE, MODE_CAPTURE, MODE_LOSS = 0, 1, 2
TAP_STYLE, TAP_CONTENT = 0, 1

c_float_p = C.POINTER(C.c_float)


class NetDesc(C.Structure):
    _fields_ = [
        ("n_entries", C.c_int),
        ("channels", C.c_int * MAUA_MAX_LAYERS),
        ("avg_pool", C.c_int),
        ("weights", C.c_void_p * MAUA_MAX_LAYERS),
        ("biases", C.c_void_p * MAUA_MAX_LAYERS),
        ("n_taps", C.c_int),
        ("tap_relu_index", C.c_int * MAUA_MAX_TAPS),
        ("tap_kind", C.c_int * MAUA_MAX_TAPS),
        ("norm_channels", C.c_int * MAUA_MAX_LAYERS),
        ("conv_kind", C.c_int * MAUA_MAX_LAYERS),
        ("pool_kind", C.c_int),
    ]


class TapIO(C.Structure):
    _fields_ = [
        ("mode", C.c_int),
        ("use_covariance", C.c_int),
        ("value_scale", C.c_float),
        ("capture_weight", C.c_float),
        ("capture_accumulate", C.c_int),
        ("target", C.c_void_p),
        ("target_elems", C.c_long),
    ]


class ImageIO(C.Structure):
    _fields_ = [
        ("tv_mode", C.c_int),
        ("tv_strength", C.c_float),
        ("temporal_mode", C.c_int),
        ("temporal_strength", C.c_float),
        ("temporal_target", C.c_void_p),
        ("temporal_target_elems", C.c_long),
        ("temporal_weights", C.c_void_p),
    ]


def declared_symbols() -> list[str]:
    """Every MAUA_API function name declared in include/maua_b200.h."""
    text = HEADER_PATH.read_text()
    return sorted(set(re.findall(r"MAUA_API\s+[\w\s\*]+?\b(maua_\w+)\s*\(", text)))


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m maua_style_b200.build` (no CPU fallback exists)")
        from . import build as _build

        _build.build()
    lib = C.CDLL(str(LIB_PATH))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    if missing and os.environ.get("MAUA_DEV_ALLOW_MISSING") == "1":  # bring-up only
        for s in missing:
            setattr(lib, s, lib.maua_abi_version)
    elif missing:
        raise RuntimeError(f"{LIB_PATH} does not export {missing}: stale build? run `python -m maua_style_b200.build --force`")
    lib.maua_last_error.restype = C.c_char_p
    lib.maua_gram_workspace_bytes.restype = C.c_size_t
    lib.maua_reduce_workspace_bytes.restype = C.c_size_t
    lib.maua_conv_first_dgrad_workspace_bytes.restype = C.c_size_t
    lib.maua_plan_device_bytes.restype = C.c_size_t
    lib.maua_plan_destroy.restype = None
    lib.maua_plan_profile_json.restype = C.c_long
    lib.maua_lbfgs_destroy.restype = None
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().maua_last_error().decode(errors="replace")
        raise RuntimeError(f"libmaua_b200 {what} failed (status {rc}): {msg}")


def ptr(t) -> C.c_void_p:
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def stream_ptr() -> C.c_void_p:
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_gpu() -> None:
    """Fail loudly when the CUDA path cannot run (no silent CPU fallback)."""
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("maua_style_b200 needs an sm_100 (B200) GPU: CUDA is not available and there is no CPU fallback")
    lib = load()
    check(lib.maua_device_check(C.c_int(torch.cuda.current_device())), "maua_device_check")
