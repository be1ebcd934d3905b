"""CPU-side checks of the C-ABI boundary: the shared library builds / loads, exports every symbol that
include/maua_b200.h declares, and refuses to compute without an sm_100 GPU (no CPU fallback)."""
import ctypes as C
import re

import pytest
import torch

from maua_style_b200 import _lib


def test_header_declares_the_expected_surface():
    syms = _lib.declared_symbols()
    for must in ["maua_conv3x3_fwd", "maua_conv3x3_dgrad", "maua_gram", "maua_style_loss_fwd", "maua_content_loss_fwd",
                 "maua_tv_loss_fwd", "maua_adam_step", "maua_lbfgs_step", "maua_plan_create", "maua_plan_forward",
                 "maua_plan_backward", "maua_last_error"]:
        assert must in syms
    text = _lib.HEADER_PATH.read_text()
    # every entry point cites the reference code it replaces somewhere in its section
    assert len(re.findall(r"(loss|models|optim)\.py:\d+", text)) >= 15


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    for s in _lib.declared_symbols():
        assert hasattr(lib, s), s
    assert lib.maua_abi_version() == 1
    assert lib.maua_reduce_workspace_bytes() > 0
    assert lib.maua_gram_workspace_bytes(512) > 512 * 512 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu():
    lib = _lib.load()
    assert lib.maua_device_check(0) != 0
    assert b"no CPU fallback" in lib.maua_last_error() or b"CUDA" in lib.maua_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.require_gpu()
    # a compute entry point must refuse too
    rc = lib.maua_adam_step(C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), C.c_long(4), C.c_float(1),
                            C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8), 1, C.c_void_p(0))
    assert rc != 0
    from maua_style_b200 import models

    class A:  # minimal args
        model_file = "nin"
        pooling = "max"
    with pytest.raises(ValueError):
        models.select_model("resnet50.pth", "max", False, False)
