"""CPU emulation of the CUDA path's arithmetic, for tolerance budgeting.  TEST INFRASTRUCTURE (see oracle/maua_oracle.py).

The sm_100a kernels differ from the fp32 oracle in exactly three places:
  * the tcgen05 MMAs read TF32 operands: conv weights are rounded to TF32 once, activations and gradients are stored
    TF32-rounded (`cvt.rna.tf32.f32`, 10-bit mantissa, ties away from zero) by the producing epilogue;
  * accumulation is fp32 (TMEM) -- like the oracle's, up to summation order;
  * conv1_1 runs in fp32 FFMA on the unrounded image; only its output is rounded.
`TF32Net` is `OracleNet` with those roundings inserted (forward: after every ReLU; backward: on the gradient leaving
every ReLU, which is where the dgrad epilogues round).  It predicts the size of the CUDA path's deviation from the reference
on any input without a GPU -- tests/test_tolerance_budget.py checks that the tolerances of the GPU parity tests are the
ones this model implies (and the measured GPU numbers in profiles/ sit where it says).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import maua_oracle as O


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32: round to a 10-bit mantissa, nearest, ties away from zero (sign-magnitude add on the bit pattern)."""
    i = x.detach().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class _RoundBothWays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return tf32_round(x)

    @staticmethod
    def backward(ctx, g):
        return tf32_round(g)


class TF32Net(O.OracleNet):
    """OracleNet with the CUDA path's operand roundings."""

    def __init__(self, params, cfg, channels=O.VGG19_CHANNELS):
        super().__init__(params, cfg, channels)
        self.rounded_weights = [tf32_round(w) for w, _ in params]

    def __call__(self, x, taps=None):
        relu_i = 0
        for kind, payload in self.seq:
            if kind == "conv":
                w, b = self.params[payload]
                x = F.conv2d(x, w if payload == 0 else self.rounded_weights[payload], b, padding=1)
            elif kind == "relu":
                x = _RoundBothWays.apply(F.relu(x))
                if taps is not None:
                    taps[self.relu_names[relu_i]] = x
                relu_i += 1
            elif kind == "pool":
                x = F.max_pool2d(x, 2, 2) if self.cfg.pooling == "max" else F.avg_pool2d(x, 2, 2)
                if self.cfg.pooling == "avg":
                    x = _RoundBothWays.apply(x)  # pool_fwd_kernel re-rounds the average
            else:
                payload.apply(x)
        return x
