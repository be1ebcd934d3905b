"""Loss modules with the reference's interface (reference loss.py), backed by the sm_100a kernels.

Same class names, constructor arguments, attributes (`mode`, `loss`, `target`, `strength`, `blend_weight`,
`weights`, `name`, `normalize`, `use_covariance`, `video_style_factor`, `reset_targets()`) and mode protocol
("none" / "capture" / "loss") as reference loss.py:32-64 (ContentLoss), :67-91 (GramMatrix), :94-186 (StyleLoss),
:224-233 (TVLoss), :10-20 (ScaleGradients).

Inside a `models.B200Net` the modules are descriptors: the network reads their mode / strength / target and runs
the fused plan (one SYRK per style layer, loss gradients folded into the dgrad kernels).  Called on their own
(`module(feature_map)`) they run the same kernels one at a time through the per-kernel C ABI, so they remain
usable as drop-in nn.Modules (batch size 1; the B > 1 window semantics of img_vid live in the network path,
maua_style_b200/window.py).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib


class ScaleGradients(torch.autograd.Function):
    """loss.py:10-20 -- identity forward; backward grad / (||grad|| + 1e-8) * strength^2.  Kept for API parity;
    the fused path applies the same rule to the scalar loss gradients on the device (maua_loss_grad_coefs)."""

    @staticmethod
    def forward(ctx, input_tensor, strength):
        ctx.strength = strength
        return input_tensor

    @staticmethod
    def backward(ctx, grad_output):
        grad_input = grad_output / (torch.norm(grad_output, keepdim=True) + 1e-8)
        return grad_input * ctx.strength * ctx.strength, None


def _to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """[1,C,H,W] CUDA tensor -> contiguous NHWC memory, TF32-rounded (kernel), returned as [H*W, C]."""
    _lib.require_gpu()
    lib = _lib.load()
    x = x.detach().float().contiguous()
    b, c, h, w = x.shape
    out = torch.empty(b, h, w, c, device=x.device)
    _lib.check(lib.maua_nchw_to_nhwc(_lib.ptr(x), _lib.ptr(out), b, c, h, w, 1, _lib.stream_ptr()), "maua_nchw_to_nhwc")
    return out


def _gram_normalised(x: torch.Tensor, use_covariance: bool):
    """gram(x) / x.nelement() for a [1,C,H,W] CUDA tensor via the tcgen05 SYRK; returns (gram [C,C], nhwc, mean)."""
    lib = _lib.load()
    b, c, h, w = x.shape
    if b != 1:
        raise NotImplementedError("maua_style_b200 supports batch size 1 only (B > 1 video-style Grams are out of scope)")
    f = _to_nhwc(x)
    ws = torch.empty(lib.maua_gram_workspace_bytes(c), dtype=torch.uint8, device=x.device)
    gram = torch.empty(c, c, device=x.device)
    mean = torch.empty(c, device=x.device)
    _lib.check(lib.maua_gram(_lib.ptr(f), C.c_long(h * w), c, int(use_covariance), _lib.ptr(gram), _lib.ptr(mean),
                             _lib.ptr(ws), _lib.MAUA_IMPL_TC, _lib.stream_ptr()), "maua_gram")
    return gram, f, mean


class GramMatrix(nn.Module):
    """loss.py:67-91: [B,C,H,W] -> un-normalised [B*C, B*C] Gram (or covariance) matrix, B = 1."""

    def forward(self, x, shift_x=0, shift_y=0, shift_t=0, flip_h=False, flip_v=False, use_covariance=False):
        if shift_x or shift_y or shift_t or flip_h or flip_v:
            raise NotImplementedError("shifted / flipped Grams are dead code in the reference (loss.py:188-221)")
        gram, _, _ = _gram_normalised(x.cuda(), use_covariance)
        return gram * float(x[0].nelement())


class _StyleLossFn(torch.autograd.Function):
    """mean((gram/nel - target)^2) for a standalone StyleLoss call; backward 4 (G-A) X / (C^3 N) via the aux GEMM."""

    @staticmethod
    def forward(ctx, x, target, use_covariance):
        lib = _lib.load()
        gram, f, mean = _gram_normalised(x, use_covariance)
        c = gram.shape[0]
        ws = torch.zeros(lib.maua_reduce_workspace_bytes(), dtype=torch.uint8, device=x.device)
        loss = torch.zeros(1, device=x.device)
        diff = torch.empty_like(gram)
        tgt = target.detach().float().contiguous()
        _lib.check(lib.maua_style_loss_fwd(_lib.ptr(gram), _lib.ptr(tgt), c, C.c_float(1.0), _lib.ptr(loss), _lib.ptr(diff),
                                           _lib.ptr(ws), _lib.stream_ptr()), "maua_style_loss_fwd")
        ctx.save_for_backward(f, diff, mean)
        ctx.use_covariance = use_covariance
        ctx.shape = x.shape
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        f, diff, mean = ctx.saved_tensors
        _, c, h, w = ctx.shape
        coef = g.detach().float().reshape(1).contiguous()
        aux_d = torch.empty_like(diff)
        aux_bias = torch.empty(c, device=f.device) if ctx.use_covariance else None
        _lib.check(lib.maua_style_loss_bwd_prep(_lib.ptr(diff), _lib.ptr(mean if ctx.use_covariance else None), c,
                                                C.c_long(h * w), _lib.ptr(coef), _lib.ptr(aux_d), _lib.ptr(aux_bias),
                                                _lib.stream_ptr()), "maua_style_loss_bwd_prep")
        gx = torch.empty_like(f)
        null = C.c_void_p(0)
        _lib.check(lib.maua_conv3x3_dgrad(null, null, _lib.ptr(gx), 1, h, w, c, c, null, _lib.ptr(f), _lib.ptr(aux_d),
                                          _lib.ptr(aux_bias), null, null, null, 0, _lib.MAUA_IMPL_TC, _lib.stream_ptr()),
                   "maua_conv3x3_dgrad(aux)")
        return gx.permute(0, 3, 1, 2), None, None


def _reduce_scratch(device) -> torch.Tensor:
    return torch.zeros(_lib.load().maua_reduce_workspace_bytes(), dtype=torch.uint8, device=device)


class _WeightedMSEFn(torch.autograd.Function):
    """mean((x * weights - target)^2) of one [1,C,H,W] frame (loss.py:53-57) through maua_content_loss_fwd; `weights` is None or a
    [1,1,H,W] map broadcast over the channels (the flow-reliability weights of the temporal loss).  Backward:
    2 weights (x weights - target) / numel."""

    @staticmethod
    def forward(ctx, x, weights, target):
        _lib.require_gpu()
        lib = _lib.load()
        xd = x.detach().cuda().float().contiguous()
        td = target.detach().to(xd.device).float().contiguous()
        wd = None
        plane = 0
        if weights is not None:
            plane = int(xd.shape[2] * xd.shape[3])
            if weights.numel() != plane:
                raise ValueError("weights must have H*W elements ([1,1,H,W])")
            wd = weights.detach().to(xd.device).float().contiguous()
        loss = torch.zeros(1, device=xd.device)
        with torch.cuda.device(xd.device):
            _lib.check(lib.maua_content_loss_fwd(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(td), C.c_long(xd.numel()), C.c_long(plane),
                                                 C.c_float(1.0), _lib.ptr(loss), _lib.ptr(_reduce_scratch(xd.device)),
                                                 _lib.stream_ptr()), "maua_content_loss_fwd")
        ctx.save_for_backward(xd, td, wd if wd is not None else torch.empty(0, device=xd.device))
        ctx.x_device = x.device
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        xd, td, wd = ctx.saved_tensors
        if wd.numel():
            w = wd.reshape(1, 1, xd.shape[2], xd.shape[3])
            gx = (xd * w - td) * w
        else:
            gx = xd - td
        return (gx * (g * (2.0 / xd.numel()))).to(ctx.x_device), None, None


class _TVFn(torch.autograd.Function):
    """strength * (sum |x[r+1] - x[r]| + sum |x[c+1] - x[c]|) (loss.py:229-233) through maua_tv_loss_fwd; backward: the signs
    of the differences scattered to both pixels (sign(0) = 0, like torch.abs)."""

    @staticmethod
    def forward(ctx, x, strength):
        _lib.require_gpu()
        lib = _lib.load()
        xd = x.detach().cuda().float().contiguous()
        b, c, h, w = xd.shape
        loss = torch.zeros(1, device=xd.device)
        with torch.cuda.device(xd.device):
            _lib.check(lib.maua_tv_loss_fwd(_lib.ptr(xd), int(b * c), int(h), int(w), C.c_float(strength), _lib.ptr(loss),
                                            _lib.ptr(_reduce_scratch(xd.device)), _lib.stream_ptr()), "maua_tv_loss_fwd")
        ctx.save_for_backward(xd)
        ctx.strength, ctx.x_device = strength, x.device
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (xd,) = ctx.saved_tensors
        gx = torch.zeros_like(xd)
        sv = torch.sign(xd[:, :, 1:, :] - xd[:, :, :-1, :])
        sh = torch.sign(xd[:, :, :, 1:] - xd[:, :, :, :-1])
        gx[:, :, 1:, :] += sv
        gx[:, :, :-1, :] -= sv
        gx[:, :, :, 1:] += sh
        gx[:, :, :, :-1] -= sh
        return (gx * (g * ctx.strength)).to(ctx.x_device), None


class ContentLoss(nn.Module):
    """loss.py:32-64."""

    def __init__(self, strength, normalize=False):
        super().__init__()
        self.strength = strength
        self.crit = nn.MSELoss()
        self.mode = "none"
        self.weights = None
        self.normalize = normalize
        self.loss = 0
        self.target = torch.Tensor()
        self.name = "cont"

    def forward(self, input):
        # standalone use (outside a B200Net, which reads the attributes and fuses the loss into the conv epilogue): the value
        # comes from the library's reduction kernel, the mode protocol is loss.py:42-64
        if self.mode == "none" or (input.shape[1:] != self.target.shape[1:] and self.target.nelement() != 0):
            return input
        if "temporal" in self.name and self.target.shape[0] == 0 and self.mode == "loss":
            return input
        if self.mode == "capture":
            self.loss = 0
            self.target = input.detach()
        elif self.mode == "loss":
            if self.target.shape[0] != 1:
                raise NotImplementedError("maua_style_b200: a standalone ContentLoss compares against a one-frame target")
            frames = input.shape[0]
            total = 0
            for idx in range(frames):
                mse = _WeightedMSEFn.apply(input[[idx]], self.weights, self.target)
                if self.normalize:
                    mse = ScaleGradients.apply(mse, self.strength)
                total = total + mse * (self.strength / frames)
            self.loss = total
        return input


class StyleLoss(nn.Module):
    """loss.py:94-186 (static + dynamic term; shift / flip losses are dead code there and absent here)."""

    def __init__(self, strength, use_covariance=False, normalize=False, video_style_factor=0, shift_factor=0,
                 flip_factor=0, rotation_factor=0):
        super().__init__()
        self.reset_targets()
        self.strength = strength
        self.blend_weight = None
        self.video_style_factor = video_style_factor
        self.shift_factor = shift_factor
        self.flip_factor = flip_factor
        self.rotation_factor = rotation_factor
        self.gram = GramMatrix()
        self.crit = nn.MSELoss()
        self.loss = 0
        self.mode = "none"
        self.use_covariance = use_covariance
        self.normalize = normalize
        self.name = "style"

    def reset_targets(self):
        self.target = torch.Tensor()
        self.video_target = torch.Tensor()
        self.shift_targets_x = []
        self.shift_targets_y = []

    def forward(self, input):
        # standalone use: one SYRK serves the static and the dynamic term (identical for B = 1)
        if self.mode == "none":
            return input
        if input.shape[0] != 1:
            raise NotImplementedError("maua_style_b200 supports batch size 1 only")
        x = input.cuda()
        if self.mode == "capture":
            gram, _, _ = _gram_normalised(x, self.use_covariance)
            self.loss = 0
            if self.target.nelement() == 0:
                self.target = self.blend_weight * gram
            else:
                self.target = self.target + self.blend_weight * gram
            self.video_target = self.target
        elif self.mode == "loss":
            mse = _StyleLossFn.apply(x, self.target, self.use_covariance)
            terms = [1.0] + ([float(self.video_style_factor)] if self.video_style_factor > 0 else [])
            for scale in terms:
                l = ScaleGradients.apply(mse, self.strength) if self.normalize else mse
                self.loss = self.loss + scale * l * self.strength
        return input


class TVLoss(nn.Module):
    """loss.py:224-233."""

    def __init__(self, strength):
        super().__init__()
        self.strength = strength
        self.loss = 0
        self.mode = "none"
        self.name = "tv"

    def forward(self, input):
        # standalone use: value from the library's kernel; `.loss` is assigned on every call, also in capture passes (loss.py:233)
        self.loss = _TVFn.apply(input, float(self.strength))
        return input


def normalize_weights(content_losses, style_losses):
    """loss.py:23-28."""
    for i in content_losses:
        i.strength = i.strength / max(i.target.size())
    for i in style_losses:
        i.strength = i.strength / max(i.target.size())
