"""CPU oracle for the maua-style VGG-19 neural-style inner loop.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product (maua_style_b200/) never does and fails
loudly when its CUDA extension is missing.

It is a functional restatement, in plain fp32 torch-on-CPU ops, of the reference's algorithm for the hot path
(citations are relative to the reference repo root, JCBrouwer/maua-style @ 316c552):

    VGG-19 feature stack ........ models.py:116-132 (build_sequential), :138 (channel_list["VGG-19"]),
                                  :205-243 (vgg19_dict), :351-453 (load_model: splice + truncate)
    GramMatrix / StyleLoss ...... loss.py:67-91, :94-186   (static + dynamic term, ScaleGradients :10-20)
    ContentLoss ................. loss.py:32-64            (incl. weighted temporal form :53-54)
    TVLoss ...................... loss.py:224-233
    feval / optimize ............ optim.py:111-255          (Adam loop runs num_iters+1 steps, :240)
    Adam / L-BFGS ............... optim.py:180-196 -> torch.optim.{Adam,LBFGS} (torch==1.8.1 pinned in
                                  requirements.txt:1; source not in the reference tree, restated here from the
                                  published algorithm: Adam without amsgrad/weight decay; L-BFGS two-loop
                                  recursion without line search, history 100, tolerance -1)

The arithmetic itself (conv2d, relu, max_pool2d, mm, abs/sum) lives in the third-party dependency PyTorch;
the oracle calls the same CPU fp32 ops through torch.nn.functional and uses autograd for the backward pass,
exactly what the reference's `total_loss.backward()` (optim.py:213) does.

Pinning: the reference ships no tests or golden vectors for this path, so the oracle
is pinned differentially: tests/golden/make_golden.py runs the UNMODIFIED reference modules (imported from
/root/reference in the build container) on seeded inputs and commits their outputs under tests/golden/;
tests/test_oracle.py checks this file against those vectors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# models.py:138
VGG19_CHANNELS = [64, 64, "P", 128, 128, "P", 256, 256, 256, 256, "P", 512, 512, 512, 512, "P", 512, 512, 512, 512, "P"]
# models.py:205-243 (names of the ReLU layers, in order)
VGG19_RELU_NAMES = ["relu1_1", "relu1_2", "relu2_1", "relu2_2", "relu3_1", "relu3_2", "relu3_3", "relu3_4",
                    "relu4_1", "relu4_2", "relu4_3", "relu4_4", "relu5_1", "relu5_2", "relu5_3", "relu5_4"]
# models.py:137 and :140-203 (VGG-16: three convs in blocks 3-5)
VGG16_CHANNELS = [64, 64, "P", 128, 128, "P", 256, 256, 256, "P", 512, 512, 512, "P", 512, 512, 512, "P"]
# models.py:136 (channel-pruned VGG-16, `--model_file *prun*`, models.py:249-258; layer names = vgg16_dict)
VGG16P_CHANNELS = [24, 22, "P", 41, 51, "P", 108, 89, 111, "P", 184, 276, 228, "P", 512, 512, 512, "P"]


# models.py:74-113 (class NIN): (output channels, kernel size) per conv -- conv1 11x11 / stride 4 / no padding, conv2 5x5 / pad 2,
# conv3 / conv4 3x3 / pad 1, "cccp" layers 1x1; pools MaxPool2d / AvgPool2d((3,3), (2,2), (0,0), ceil_mode=True).  load_model keeps
# only Conv2d / ReLU / pool modules (models.py:381-436): the Dropout in front of conv4 is never part of the style network.
NIN_LAYERS = [(96, 11), (96, 1), (96, 1), "P", (256, 5), (256, 1), (256, 1), "P", (384, 3), (384, 1), (384, 1), "P",
              (1024, 3), (1024, 1), (1000, 1)]
_CONV_GEOMETRY = {3: (1, 1), 1: (1, 0), 5: (1, 2), 11: (4, 0)}  # kernel -> (stride, padding)


def is_nin(channels) -> bool:
    return any(isinstance(c, tuple) for c in channels)


def relu_names(channels) -> List[str]:
    """models.py:140-243 (vgg16_dict / vgg19_dict "R" lists): relu{block}_{index} in network order; nin_dict (:140-172): relu1.."""
    if is_nin(channels):
        return [f"relu{i + 1}" for i in range(sum(1 for c in channels if c != "P"))]
    names, block, idx = [], 1, 1
    for c in channels:
        if c == "P":
            block, idx = block + 1, 1
        else:
            names.append(f"relu{block}_{idx}")
            idx += 1
    return names


# load.py:30 -- BGR channel means subtracted from the 0-255 image
BGR_MEAN = (103.939, 116.779, 123.68)


@dataclass
class StyleConfig:
    """The subset of the reference's args Namespace that the hot path reads (config.py:12-91, defaults there)."""
    content_layers: str = "relu4_2"
    style_layers: str = "relu1_1,relu2_1,relu3_1,relu4_1,relu5_1"
    content_weight: float = 5.0
    style_weight: float = 100.0
    tv_weight: float = 1e-3
    temporal_weight: float = 50.0
    pooling: str = "max"
    use_covariance: bool = False
    normalize_gradients: bool = True
    video_style_factor: float = 100.0
    normalize_weights: bool = False
    style_blend_weights: Optional[Sequence[float]] = None  # normalised to sum 1 (config.py:146-164)
    optimizer: str = "adam"
    learning_rate: float = 1.0
    lbfgs_num_correction: int = 100

    def blend(self, n_styles: int) -> List[float]:
        w = list(self.style_blend_weights) if self.style_blend_weights is not None else [1.0] * n_styles
        s = sum(w)
        return [x / s for x in w]


def he_init_vgg19(seed: int = 0, channels=VGG19_CHANNELS) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Random-init VGG-19 conv weights (pretrained checkpoints are unavailable offline): He-normal weights,
    bias sigma 0.1, fixed seed -- SURVEY.md section 8a hazard (4)."""
    g = torch.Generator().manual_seed(seed)
    params, cin = [], 3
    for c in channels:
        if c == "P":
            continue
        c, ks = c if isinstance(c, tuple) else (c, 3)
        w = torch.randn(c, cin, ks, ks, generator=g) * math.sqrt(2.0 / (cin * ks * ks))
        b = torch.randn(c, generator=g) * 0.1
        params.append((w, b))
        cin = c
    return params


def synthetic_image(h: int, w: int, seed: int, smooth: bool = False) -> torch.Tensor:
    """U(0,255) - BGR mean, the value range load.preprocess produces (load.py:21-32)."""
    g = torch.Generator().manual_seed(seed)
    if smooth:
        low = torch.rand(1, 3, max(h // 8, 2), max(w // 8, 2), generator=g)
        img = F.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)
    else:
        img = torch.rand(1, 3, h, w, generator=g)
    return img * 255.0 - torch.tensor(BGR_MEAN).view(1, 3, 1, 1)


# ------------------------------------------------------------------------------------------------------------
# loss math
# ------------------------------------------------------------------------------------------------------------
class _ScaleGradients(torch.autograd.Function):
    """loss.py:10-20: identity forward; backward grad / (||grad|| + 1e-8) * strength^2."""

    @staticmethod
    def forward(ctx, x, strength):
        ctx.strength = strength
        return x

    @staticmethod
    def backward(ctx, g):
        g = g / (torch.norm(g, keepdim=True) + 1e-8)
        return g * ctx.strength * ctx.strength, None


def gram_matrix(x: torch.Tensor, use_covariance: bool = False) -> torch.Tensor:
    """loss.py:67-91 with y = x: [B,C,H,W] -> [B*C, B*C]."""
    B, C, H, W = x.shape
    xf = x.reshape(B * C, H * W)
    if use_covariance:
        xf = xf - xf.mean(1).unsqueeze(1)
    return torch.mm(xf, xf.t())


@dataclass
class LossModule:
    kind: str                     # "style" | "content" | "tv" | "temporal"
    strength: float
    name: str = ""
    mode: str = "none"            # "none" | "capture" | "loss"
    normalize: bool = False
    use_covariance: bool = False
    video_style_factor: float = 0.0
    blend_weight: Optional[float] = None
    target: torch.Tensor = field(default_factory=torch.Tensor)
    video_target: torch.Tensor = field(default_factory=torch.Tensor)
    weights: Optional[torch.Tensor] = None
    loss: object = 0

    def reset_targets(self):  # loss.py:124-128
        self.target = torch.Tensor()
        self.video_target = torch.Tensor()

    # -- forward of each module kind -------------------------------------------------------------
    def apply(self, x: torch.Tensor) -> None:
        if self.kind == "tv":  # loss.py:229-233 (assigned on every forward, any mode)
            xd = x[:, :, 1:, :] - x[:, :, :-1, :]
            yd = x[:, :, :, 1:] - x[:, :, :, :-1]
            self.loss = self.strength * (torch.sum(torch.abs(xd)) + torch.sum(torch.abs(yd)))
        elif self.kind in ("content", "temporal"):
            self._content(x)
        elif self.kind == "style":
            if self.mode == "none":
                return
            self._style_static(x)
            if self.video_style_factor > 0:
                self._style_dynamic(x)

    def _content(self, x):  # loss.py:42-64
        if self.mode == "none" or (x.shape[1:] != self.target.shape[1:] and self.target.nelement() != 0):
            return
        if self.kind == "temporal" and self.target.shape[0] == 0 and self.mode == "loss":
            return
        self.loss = 0
        for idx in range(x.shape[0]):
            if self.mode == "loss":
                xi = x[[idx]]
                l = F.mse_loss(xi * self.weights, self.target) if self.weights is not None else F.mse_loss(xi, self.target)
                if self.normalize:
                    l = _ScaleGradients.apply(l, self.strength)
                self.loss = self.loss + l * self.strength / x.shape[0]
            if self.mode == "capture":
                self.target = x.detach()

    def _style_static(self, x):  # loss.py:141-157
        for idx in range(x.shape[0]):
            g = gram_matrix(x[idx].unsqueeze(0), self.use_covariance) / x[idx].nelement()
            if self.mode == "capture":
                self.loss = 0
                if self.target.nelement() == 0:
                    self.target = self.blend_weight * g.detach() / x.shape[0]
                else:
                    self.target = self.target + self.blend_weight * g.detach() / x.shape[0]
            if self.mode == "loss":
                l = F.mse_loss(g, self.target)
                if self.normalize:
                    l = _ScaleGradients.apply(l, self.strength)
                self.loss = self.loss + l * self.strength / x.shape[0]

    def _style_dynamic(self, x):  # loss.py:164-181
        if self.video_target.nelement() != 0 and gram_matrix(x, False).shape[0] != self.video_target.shape[0]:
            return
        g = gram_matrix(x, self.use_covariance) / x.nelement()
        if self.mode == "capture":
            self.loss = 0
            if self.video_target.nelement() == 0:
                self.video_target = self.blend_weight * g.detach()
            else:
                self.video_target = self.video_target + self.blend_weight * g.detach()
        if self.mode == "loss":
            l = F.mse_loss(g, self.video_target)
            if self.normalize:
                l = _ScaleGradients.apply(l, self.strength)
            self.loss = self.loss + self.video_style_factor * l * self.strength / x.shape[0]


# ------------------------------------------------------------------------------------------------------------
# network
# ------------------------------------------------------------------------------------------------------------
class OracleNet:
    """models.py:351-453: TVLoss, temporal ContentLoss, then conv/relu/pool with loss modules spliced after the
    named ReLUs; the stack stops after the last requested tap (models.py:382)."""

    def __init__(self, params, cfg: StyleConfig, channels=VGG19_CHANNELS):
        self.cfg = cfg
        self.params = params
        self.channels = channels
        self.nin = is_nin(channels)
        self.relu_names = relu_names(channels)
        content_layers = cfg.content_layers.split(",")
        style_layers = cfg.style_layers.split(",")
        self.tv_losses, self.temporal_losses, self.content_losses, self.style_losses = [], [], [], []
        self.seq: list = []  # ("tv"|"temporal"|"conv"|"relu"|"pool"|"loss", payload)
        if cfg.tv_weight > 0:
            m = LossModule("tv", cfg.tv_weight, name=f"tv {len(self.seq)}")
            self.seq.append(("loss", m)); self.tv_losses.append(m)
        if cfg.temporal_weight > 0:
            m = LossModule("temporal", cfg.temporal_weight, name=f"temporal {len(self.seq)}", normalize=cfg.normalize_gradients)
            self.seq.append(("loss", m)); self.temporal_losses.append(m)
        next_c, next_s, conv_i, relu_i = 1, 1, 0, 0
        for c in channels:
            if not (next_c <= len(content_layers) or next_s <= len(style_layers)):
                break
            if c == "P":
                self.seq.append(("pool", None))
                continue
            self.seq.append(("conv", conv_i)); conv_i += 1
            self.seq.append(("relu", None))
            name = self.relu_names[relu_i]; relu_i += 1
            if name in content_layers:
                m = LossModule("content", cfg.content_weight, name=f"cont {len(self.seq)}", normalize=cfg.normalize_gradients)
                self.seq.append(("loss", m)); self.content_losses.append(m); next_c += 1
            if name in style_layers:
                m = LossModule("style", cfg.style_weight, name=f"style {len(self.seq)}", normalize=cfg.normalize_gradients,
                               use_covariance=cfg.use_covariance, video_style_factor=cfg.video_style_factor)
                self.seq.append(("loss", m)); self.style_losses.append(m); next_s += 1
        self.losses = self.content_losses + self.style_losses + self.tv_losses + self.temporal_losses  # models.py:453

    def __call__(self, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        relu_i = 0
        for kind, payload in self.seq:
            if kind == "conv":
                w, b = self.params[payload]
                stride, pad = _CONV_GEOMETRY[w.shape[-1]]
                x = F.conv2d(x, w, b, stride=stride, padding=pad)
            elif kind == "relu":
                x = F.relu(x)
                if taps is not None:
                    taps[self.relu_names[relu_i]] = x
                relu_i += 1
            elif kind == "pool" and self.nin:  # models.py:77-80
                x = (F.max_pool2d(x, 3, 2, 0, ceil_mode=True) if self.cfg.pooling == "max"
                     else F.avg_pool2d(x, 3, 2, 0, ceil_mode=True))
            elif kind == "pool":
                x = F.max_pool2d(x, 2, 2) if self.cfg.pooling == "max" else F.avg_pool2d(x, 2, 2)
            else:
                payload.apply(x)
        return x


# optim.py:22-66
def set_content_targets(net: OracleNet, content: torch.Tensor):
    for m in net.content_losses:
        m.mode = "capture"
    net(content)
    for m in net.content_losses:
        m.mode = "none"


def set_temporal_targets(net: OracleNet, warp: torch.Tensor, weights: Optional[torch.Tensor] = None):
    for m in net.temporal_losses:
        m.mode = "capture"
        if weights is not None:
            m.weights = weights
    net(warp)
    for m in net.temporal_losses:
        m.mode = "none"


def set_style_targets(net: OracleNet, styles: Sequence[torch.Tensor], blend: Sequence[float]):
    for m in net.style_losses:
        m.reset_targets()
        m.mode = "capture"
    for i, img in enumerate(styles):
        for m in net.style_losses:
            m.blend_weight = blend[i]
        net(img)
    for m in net.style_losses:
        m.mode = "none"


def feval(net: OracleNet, pastiche: torch.Tensor):
    """optim.py:201-238: returns (total loss, per-module loss values, d total / d pastiche)."""
    x = pastiche.detach().clone().requires_grad_(True)
    net(x)
    total, values = 0, []
    for m in net.losses:
        if isinstance(m.loss, int) and m.loss == 0:
            values.append(0.0)
            continue
        values.append(float(m.loss.detach()))
        total = total + m.loss
    total.backward()
    for m in net.losses:
        m.loss = 0
    return float(total.detach()), values, x.grad.detach()


# ------------------------------------------------------------------------------------------------------------
# optimizers (restated; torch.optim is NOT used here)
# ------------------------------------------------------------------------------------------------------------
def adam_optimize(p: torch.Tensor, closure: Callable, steps: int, lr: float = 1.0, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (no amsgrad / weight decay) on one tensor; closure(p) -> grad."""
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for t in range(1, steps + 1):
        g = closure(p)
        m = m + (g - m) * (1 - b1)
        v = v * b2 + (1 - b2) * g * g
        bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
        p = p - (lr / bc1) * (m / (v.sqrt() / math.sqrt(bc2) + eps))
    return p


def lbfgs_optimize(p: torch.Tensor, closure: Callable, max_iter: int, lr: float = 1.0, history: int = 100,
                   tolerance_change: float = -1.0):
    """torch.optim.LBFGS.step without line search, tolerance_grad = -1 (never satisfied); one call = max_iter
    closure evaluations and max_iter parameter updates (optim.py:180-191, :240)."""
    # torch.optim.LBFGS also stops on max_eval = max_iter * 5 // 4 closure evaluations (its default; optim.py:180-191 does not
    # pass it): one evaluation before the loop + one per iteration, so max_iter = 2 / 3 give only 1 / 2 updates
    max_iter = min(max_iter, max(max_iter * 5 // 4 - 1, 1))
    p = p.clone()
    flat = p.view(-1)
    g = closure(p).reshape(-1)
    old_dirs, old_stps, ro = [], [], []
    H_diag, d, t, prev_g = 1.0, None, None, None
    n_iter = 0
    while n_iter < max_iter:
        n_iter += 1
        if n_iter == 1:
            d = g.neg()
        else:
            y = g.sub(prev_g)
            s = d.mul(t)
            ys = float(y.dot(s))
            if ys > 1e-10:
                if len(old_dirs) == history:
                    old_dirs.pop(0); old_stps.pop(0); ro.pop(0)
                old_dirs.append(y); old_stps.append(s); ro.append(1.0 / ys)
                H_diag = ys / float(y.dot(y))
            k = len(old_dirs)
            al = [0.0] * k
            q = g.neg()
            for i in range(k - 1, -1, -1):
                al[i] = float(old_stps[i].dot(q)) * ro[i]
                q.add_(old_dirs[i], alpha=-al[i])
            d = r = q * H_diag
            for i in range(k):
                be = float(old_dirs[i].dot(r)) * ro[i]
                r.add_(old_stps[i], alpha=al[i] - be)
        prev_g = g.clone()
        g1 = float(g.abs().sum())  # (torch divides tensors: 1 / 0 = inf and min(1, inf) = 1 -- an all-masked img_vid window)
        t = min(1.0, 1.0 / g1 if g1 > 0 else float("inf")) * lr if n_iter == 1 else lr
        gtd = float(g.dot(d))
        if gtd > -tolerance_change:
            break
        flat.add_(d, alpha=t)
        if n_iter != max_iter:
            g = closure(p).reshape(-1)
    return p


def optimize(content, styles, init, num_iters, cfg: StyleConfig, params, temporal=None, channels=VGG19_CHANNELS):
    """optim.py:111-255 for transfer_type img_img (one window): capture targets, then Adam (num_iters + 1 steps,
    the reference's off-by-one at :240) or one L-BFGS step() of num_iters iterations."""
    net = OracleNet(params, cfg, channels)
    set_content_targets(net, content)
    if temporal is not None:
        set_temporal_targets(net, *temporal)
    set_style_targets(net, styles, cfg.blend(len(styles)))
    for m in net.losses:
        m.mode = "loss"
    if cfg.normalize_weights:  # optim.py:176-178
        for m in net.content_losses + net.style_losses + net.temporal_losses:
            m.strength = m.strength / max(m.target.size())

    def closure(p):
        return feval(net, p)[2]

    p0 = init.clone().float()
    if cfg.optimizer == "adam":
        return adam_optimize(p0, closure, num_iters + 1, lr=cfg.learning_rate)
    return lbfgs_optimize(p0, closure, num_iters, lr=1.0, history=cfg.lbfgs_num_correction)


# ------------------------------------------------------------------------------------------------------------
# transfer type img_vid: windows of B > 1 frames (optim.py:69-90, :113-125, :149-170, :216-219, :242-247)
# ------------------------------------------------------------------------------------------------------------
def set_style_video_targets(net: OracleNet, videos: Sequence[torch.Tensor], blend: Sequence[float], gfw: int):
    """optim.py:69-90: every style clip is averaged over all of its windows of `gfw` frames."""
    for m in net.style_losses:
        m.reset_targets()
        m.mode = "capture"
    for i, video in enumerate(videos):
        n_win = max(len(video) - gfw + 1, 1)
        for m in net.style_losses:
            m.blend_weight = blend[i] / n_win
        for start in range(n_win):
            net(video[start:start + gfw])
    for m in net.style_losses:
        m.mode = "none"


def wrapping_indices(n: int, start: int, length: int) -> torch.Tensor:
    """utils.py:76-85."""
    if start + length <= n:
        idx = torch.arange(start, start + length)
    else:
        idx = torch.cat((torch.arange(start, n), torch.arange(0, (start + length) % n)))
    if n == 1:
        idx = torch.zeros(1, dtype=torch.int64)
    return idx


def optimize_windows(content, styles, init, num_iters, cfg: StyleConfig, params, gfw: int, afw: int = -1,
                     channels=VGG19_CHANNELS):
    """optim.py:111-255 with "_vid" in transfer_type: the pastiche video is optimised window by window; frames the previous
    window styled get zero gradient (:216-219); a fresh optimizer per window."""
    clips = [init] + list(styles)
    num_windows = math.ceil(init.shape[0] / gfw)
    framestep = [(c.shape[0] - gfw / 2) / num_windows for c in clips]
    windows = [[math.ceil(framestep[i] * n) for n in range(num_windows + 1)] if clips[i].shape[0] != 1
               else [0] * (num_windows + 1) for i in range(len(clips))]
    net = OracleNet(params, cfg, channels)
    blend = cfg.blend(len(styles))
    set_content_targets(net, content)
    if afw == -1:
        set_style_video_targets(net, styles, blend, gfw)
        for m in net.losses:
            m.mode = "loss"
    output = init.clone().float()
    T = output.shape[0]
    for w, start in enumerate(windows[0]):
        front = windows[0][w - 1] + gfw - start
        end = (start + gfw) % T if start + gfw >= T else 0
        idx = wrapping_indices(T, start, gfw)
        if afw != -1:
            cur = [s[wrapping_indices(s.shape[0], windows[k + 1][w], afw)] for k, s in enumerate(styles)]
            set_style_video_targets(net, cur, blend, gfw)
            for m in net.losses:
                m.mode = "loss"
        if w == 0 and cfg.normalize_weights:
            for m in net.content_losses + net.style_losses + net.temporal_losses:
                m.strength = m.strength / max(m.target.size())

        def closure(p, w=w, front=front, end=end):
            g = feval(net, p)[2]
            if w != 0:
                g[:front] = 0
                if end > 0:
                    g[-end:] = 0
            return g

        p0 = output[idx].clone()
        if cfg.optimizer == "adam":
            res = adam_optimize(p0, closure, num_iters + 1, lr=cfg.learning_rate)
        else:
            res = lbfgs_optimize(p0, closure, num_iters, lr=1.0, history=cfg.lbfgs_num_correction)
        output[idx] = res
    return output


def psnr(a: torch.Tensor, b: torch.Tensor, peak: float = 255.0) -> float:
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)
