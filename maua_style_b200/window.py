"""`net(x)` for a window of B > 1 frames -- the batch semantics of the reference's loss modules that `img_vid` uses
(reference loss.py:42-64, :141-181; optim.py:113-125, :149-170, :216-219).

What the reference computes for x = [B, 3, H, W] (SURVEY.md section 8a, R4/R5/R7 with B > 1):

  ContentLoss ........ sum_b strength * MSE(x_b, target) / B            (target captured from ONE content image)
  StyleLoss static ... sum_b strength * MSE(G_b, A) / B,  G_b = Gram(x_b) / (C H W)     -- per-frame [C, C] Grams
  StyleLoss dynamic .. vsf * strength * MSE(G, A_v) / B,  G = Gram(x) / (B C H W)       -- ONE [B C, B C] Gram over the
                       channels of all frames; skipped when the captured video target is not [B C, B C] (image styles)
  TVLoss ............. strength * (sum |dx| + sum |dy|) over the whole batch
  ScaleGradients ..... per term: the gradient of every per-frame / dynamic MSE is scaled to strength^2

How it runs here.  Every frame goes through its own plan (same weights, own activation arena), so the per-frame
feature stacks are the single-image tcgen05 path unchanged.  At a style tap the B feature maps are laid side by side as
X = [H_l W_l, B C] (maua_plan_tap_feature_strided); ONE SYRK over X (maua_gram on B C channels -- the same Gram kernel)
gives the dynamic Gram, and its diagonal [C, C] blocks are the B static Grams (x B, the normalisations differ by the
factor B).  The backward term of frame b is one GEMM row block,

    d loss / d F_b = X . M[bC:(b+1)C, :]^T,      M = c_d k_d (G - A_v)  +  blockdiag_b( c_s k_s (G_b - A) ),

which maua_plan_set_tap_fold folds into that frame's dgrad launch as B C / 32 extra k-steps of the tensor-core GEMM (the
same mechanism that folds the B = 1 StyleLoss backward).  The [B C, B C] bookkeeping between the two (differences, MSE
values, the coefficient matrix M) is a handful of small element-wise device ops.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib
from .loss import ContentLoss, StyleLoss

MODE_EXTERNAL = 3


def _sg(x: torch.Tensor) -> torch.Tensor:
    """ScaleGradients on a scalar (loss.py:10-20): grad / (|grad| + 1e-8)."""
    return x / (x.abs() + 1e-8)


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32 on a tensor (round to nearest, ties away, 10 mantissa bits): the tensor core reads the top 19 bits
    of an fp32 operand, so matrices handed to it as GEMM operands are rounded first (csrc/common.cuh round_tf32)."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class FrameWindow:
    def __init__(self, net):
        self.net = net
        self.lib = net._lib
        self.cores = []          # extra plan cores for frames 1..B-1 (frame 0 runs on the network's own plan)
        self.loss_vecs: List[torch.Tensor] = []
        self.coefs: List[torch.Tensor] = []
        self.xcat = {}           # tap -> [P, B*C] side-by-side tap features
        self.gram = {}           # tap -> (gram [BC, BC], mean [BC], workspace)
        self.state = {}          # tap -> tensors kept for the backward pass
        self._keep = []
        self._scale = None       # per-slot upstream scale (1/B for the per-frame means, 1 for TVLoss) on the device

    # ------------------------------------------------------------------------------------------------------------
    def _plans(self, B: int):
        from .models import _PlanCore

        net = self.net
        if net.n_stages != 1:
            raise NotImplementedError("maua_style_b200: windows of B > 1 frames are not combined with the layer-wise multidevice split")
        core = net._core
        while len(self.cores) < B - 1:
            params = [(w.data, b.data) for w, b in zip(core.weights, core.biases)]
            self.cores.append(_PlanCore(core.entries, params, core.avg_pool, core.tap_sig, net.device, core.bounds, core.devs,
                                        core.norm_channels, core.conv_kinds, core.pool_kind))
        plans = [net._plan] + [c.stages[0]["plan"] for c in self.cores[:B - 1]]
        impl = net._impl
        for c in self.cores[:B - 1]:
            if getattr(c, "_impl_set", None) != impl:  # (not on every pass: the call synchronises, which a graph capture forbids)
                _lib.check(self.lib.maua_plan_set_impl(c.stages[0]["plan"], impl), "maua_plan_set_impl")
                c._impl_set = impl
        while len(self.loss_vecs) < B:
            self.loss_vecs.append(torch.zeros(net._n_slots, device=net.device))
            self.coefs.append(torch.zeros(net._n_slots, device=net.device))
        return plans

    def _dynamic_live(self, mod: StyleLoss, B: int, C_: int) -> bool:
        """loss.py:165-166: the dynamic term is skipped when a captured video target has another size."""
        if not (float(mod.video_style_factor) > 0):
            return False
        vt = mod.video_target
        return not (vt.nelement() != 0 and vt.shape[0] != B * C_)

    # ------------------------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, keep: bool) -> None:
        net, lib = self.net, self.lib
        B, H, W = int(x.shape[0]), int(x.shape[2]), int(x.shape[3])
        plans = self._plans(B)
        for t, (ridx, mod) in enumerate(net.taps):
            if net._tap_cpad[t] != net._tap_channels[t]:
                raise NotImplementedError("maua_style_b200: windows of B > 1 frames on a zero-padded (pruned) network")
            if isinstance(mod, ContentLoss) and mod.mode == "capture":
                raise NotImplementedError("maua_style_b200: content targets are captured from one image (optim.py:22-32)")
        tm = net.temporal_mod
        if tm is not None and tm.mode != "none" and tm.target.nelement() != 0:
            raise NotImplementedError("maua_style_b200: the temporal ContentLoss is not evaluated on windows of B > 1 frames")

        # which style taps carry the dynamic term in this pass; fresh static targets are created here (loss.py:146-151)
        dyn, fresh = {}, {}
        for t, (ridx, mod) in enumerate(net.taps):
            if not isinstance(mod, StyleLoss) or mod.mode == "none":
                continue
            C_ = net._tap_channels[t]
            dyn[t] = self._dynamic_live(mod, B, C_)
            fresh[t] = mod.target.nelement() == 0
            if mod.mode == "capture":
                mod.loss = 0
                if fresh[t]:
                    mod.target = torch.zeros(C_, C_, device=net._tap_device(t))

        saved_mode = {}
        if tm is not None:
            saved_mode["tm"], tm.mode = tm.mode, "none"
        base_tio, base_iio = net._build_io(H, W, window=True)
        if tm is not None:
            tm.mode = saved_mode["tm"]
        n_t = max(len(net.taps), 1)
        self._keep = [x]
        total = None
        for b in range(B):
            tio = (_lib.TapIO * n_t)()
            C.memmove(tio, base_tio, C.sizeof(tio))
            for t, (ridx, mod) in enumerate(net.taps):
                io = tio[t]
                if isinstance(mod, StyleLoss):
                    if mod.mode == "none":
                        continue
                    if dyn[t]:
                        io.mode = MODE_EXTERNAL
                    elif mod.mode == "loss":
                        io.value_scale = float(mod.strength) / B          # loss.py:157
                    else:
                        io.capture_weight = float(mod.blend_weight) / B   # loss.py:148-151
                        io.capture_accumulate = 0 if (fresh[t] and b == 0) else 1
                else:
                    io.value_scale = float(mod.strength) / B              # loss.py:59
            xb = x[b]
            self.loss_vecs[b].zero_()  # slots this pass does not write must not carry values of an earlier pass
            self._keep += [tio, base_iio]
            with torch.cuda.device(net.device):
                _lib.check(lib.maua_plan_forward(plans[b], _lib.ptr(xb), H, W, tio, C.byref(base_iio), _lib.ptr(self.loss_vecs[b]),
                                                 int(keep), _lib.stream_ptr()), "maua_plan_forward")
            total = self.loss_vecs[b].clone() if total is None else total + self.loss_vecs[b]

        # style taps with the dynamic term: ONE SYRK over the side-by-side features of all frames
        self.state = {}
        for t, (ridx, mod) in enumerate(net.taps):
            if not dyn.get(t, False):
                continue
            C_ = net._tap_channels[t]
            h, w = net._tap_hw(H, W, ridx)
            P, BC = h * w, B * C_
            dev = net._tap_device(t)
            xc = self.xcat.get(t)
            if xc is None or tuple(xc.shape) != (P, BC):
                xc = self.xcat[t] = torch.empty(P, BC, device=dev)
            g = self.gram.get(t)
            if g is None or g[0].shape[0] != BC:
                ws = torch.empty(int(lib.maua_gram_workspace_bytes(BC)) // 4 + 64, device=dev)
                g = self.gram[t] = (torch.empty(BC, BC, device=dev), torch.empty(BC, device=dev), ws)
            gram, mean, ws = g
            cov = bool(mod.use_covariance)
            with torch.cuda.device(dev):
                for b in range(B):
                    _lib.check(lib.maua_plan_tap_feature_strided(plans[b], t, C.c_void_p(xc.data_ptr() + 4 * b * C_),
                                                                 C.c_long(BC), _lib.stream_ptr()), "maua_plan_tap_feature_strided")
                _lib.check(lib.maua_gram(_lib.ptr(xc), C.c_long(P), BC, int(cov), _lib.ptr(gram), _lib.ptr(mean), _lib.ptr(ws),
                                         net._impl, _lib.stream_ptr()), "maua_gram")
            # per-frame static Grams = diagonal blocks x B  (Gram(x_b) / (C H W) vs Gram(x) / (B C H W))
            stat = torch.diagonal(gram.view(B, C_, B, C_), dim1=0, dim2=2).permute(2, 0, 1) * float(B)  # [B, C, C]
            if mod.mode == "capture":
                bw = float(mod.blend_weight)
                mod.target += bw * stat.sum(0) / B                                    # loss.py:148-151
                if mod.video_target.nelement() == 0:                                  # loss.py:172-175
                    mod.video_target = bw * gram
                else:
                    mod.video_target = mod.video_target + bw * gram
            else:
                s, vsf = float(mod.strength), float(mod.video_style_factor)
                ds = stat - mod.target.to(dev)                                       # [B, C, C]
                dd = gram - mod.video_target.to(dev)                                  # [BC, BC]
                value = s * (ds * ds).mean(dim=(1, 2)).sum() / B + vsf * s * (dd * dd).mean() / B   # loss.py:153-157, :177-181
                total[t] = total[t] + value
                self.state[t] = (ds, dd, mean if cov else None, P, C_, BC)
        net._loss_vec.copy_(total)
        self._B, self._plans_used = B, plans

    # ------------------------------------------------------------------------------------------------------------
    def backward(self, up: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        net, lib = self.net, self.lib
        B, plans = self._B, self._plans_used
        n = net._n_slots
        strength = (C.c_float * n)()
        vsf0 = (C.c_float * n)()
        normalize = (C.c_int * n)()
        kind = (C.c_int * n)()
        scale = [1.0] * n
        for i, mod in enumerate(net.slot_modules()):
            if mod is None:
                continue
            strength[i] = float(mod.strength)
            if isinstance(mod, StyleLoss):
                kind[i], normalize[i] = 0, int(bool(mod.normalize))
                scale[i] = 1.0 / B
            elif isinstance(mod, ContentLoss):
                kind[i], normalize[i] = 1, int(bool(mod.normalize))
                scale[i] = 1.0 / B
            else:
                kind[i] = 2  # TVLoss: a sum over the batch, never normalised
        if self._scale is None or self._scale[0] != scale:
            # (made once, in the eager warm-up iterations: a host -> device copy cannot be captured into a CUDA graph)
            self._scale = (scale, torch.tensor(scale, device=net.device))
        up = up.detach().to(net.device, torch.float32)
        up_f = (up * self._scale[1]).contiguous()
        exact = net._impl == _lib.MAUA_IMPL_FP32
        keep = []
        for t, (ds, dd, mean, P, C_, BC) in self.state.items():
            mod = net.taps[t][1]
            s, vsf = float(mod.strength), float(mod.video_style_factor)
            u = up[t]
            if mod.normalize:
                c_s = _sg(u * s / B) * s * s
                c_d = _sg(u * vsf * s / B) * s * s
            else:
                c_s = u * s / B
                c_d = u * vsf * s / B
            M = (c_d * (4.0 / (float(BC) ** 3 * P))) * dd
            blocks = torch.diagonal(M.view(B, C_, B, C_), dim1=0, dim2=2)       # [C, C, B] view into M
            blocks += ((c_s * (4.0 / (float(C_) ** 3 * P))) * ds).permute(1, 2, 0)
            if not exact:
                M = round_tf32(M)
            bias = None
            if mean is not None:  # covariance: d/dF of the centred Gram = M (X - mu)  =>  bias = -M mu
                bias = -(M @ mean).contiguous()
            keep += [M, bias]
            for b in range(B):
                _lib.check(lib.maua_plan_set_tap_fold(plans[b], t, _lib.ptr(self.xcat[t]), BC, C.c_void_p(M.data_ptr() + 4 * b * C_ * BC),
                                                      C.c_void_p(bias.data_ptr() + 4 * b * C_) if bias is not None else C.c_void_p(0)),
                           "maua_plan_set_tap_fold")
        grad = torch.empty_like(x)
        with torch.cuda.device(net.device):
            for b in range(B):
                _lib.check(lib.maua_loss_grad_coefs(_lib.ptr(up_f), _lib.ptr(self.coefs[b]), n, strength, vsf0, normalize, kind,
                                                    _lib.stream_ptr()), "maua_loss_grad_coefs")
                _lib.check(lib.maua_plan_backward(plans[b], _lib.ptr(self.coefs[b]), _lib.ptr(grad[b]), _lib.stream_ptr()),
                           "maua_plan_backward")
        self._keep_bwd = keep + [up_f]
        return grad
