"""`load_model(args) -> (net, losses)` with the reference's interface (reference models.py:351-453), backed by a
maua_plan (the fused sm_100a forward / backward launch sequence in csrc/plan.cu).

What is kept from the reference:
  * the channel lists and layer-name tables (models.py:135-139, :140-243) for the VGG-16 / VGG-19 stacks;
  * checkpoint loading: `torch.load(model_file)` of a torchvision-style state dict with `features.N.weight/bias`
    keys, `disable_check` => non-strict (models.py:343);
  * module splicing order TVLoss, temporal ContentLoss, then ContentLoss / StyleLoss after the named layers, and
    truncation after the last requested tap (models.py:369-436, :382);
  * the returned `net` is callable on a [1,3,H,W] image, carries `.content_losses / .style_losses / .tv_losses /
    .temporal_losses` (models.py:447-451) and leaves 0-dim autograd-connected tensors in each module's `.loss`.
What is different: `net` is a single nn.Module (B200Net), not an nn.Sequential of torch layers; the weights live in
GEMM layout inside the plan.  Only `relu*` taps are supported -- the reference's `conv*` taps alias a buffer that the
in-place ReLU overwrites (SURVEY.md section 8a hazard 1), so they never meant what they say.
"""
from __future__ import annotations

import collections
import ctypes as C
import os
import weakref
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib
from .loss import ContentLoss, ScaleGradients, StyleLoss, TVLoss  # noqa: F401  (models.py:13 `from loss import *`)

# models.py:135-139
channel_list = {
    "VGG-16p": [24, 22, "P", 41, 51, "P", 108, 89, 111, "P", 184, 276, 228, "P", 512, 512, 512, "P"],
    "VGG-16": [64, 64, "P", 128, 128, "P", 256, 256, 256, "P", 512, 512, 512, "P", 512, 512, 512, "P"],
    "VGG-19": [64, 64, "P", 128, 128, "P", 256, 256, 256, 256, "P", 512, 512, 512, 512, "P", 512, 512, 512, 512, "P"],
}


# models.py:74-113 (class NIN): (output channels, kernel) per conv; conv1 is 11x11 / stride 4 / no padding, conv2 5x5 / pad 2,
# conv3 / conv4 3x3 / pad 1, the "cccp" layers 1x1; pools are 3x3 / stride 2 / ceil_mode.  The Dropout in front of conv4 and
# the classifier tail (6x6 average pool, Softmax) are never part of the style network: load_model only keeps Conv2d / ReLU /
# pool modules up to the last tapped layer (models.py:381-436).
NIN_LAYERS = [(96, 11), (96, 1), (96, 1), "P", (256, 5), (256, 1), (256, 1), "P", (384, 3), (384, 1), (384, 1), "P",
              (1024, 3), (1024, 1), (1000, 1)]
# index of every conv inside NIN.features (ReLU after each conv, pools at 6 / 13 / 20, Dropout at 21)
NIN_FEATURE_INDEX = [0, 2, 4, 7, 9, 11, 14, 16, 18, 22, 24, 26]
# models.py:140-172
nin_dict = {
    "C": ["conv1", "cccp1", "cccp2", "conv2", "cccp3", "cccp4", "conv3", "cccp5", "cccp6", "conv4-1024", "cccp7-1024", "cccp8-1024"],
    "R": [f"relu{i}" for i in range(1, 13)],
    "P": ["pool1", "pool2", "pool3", "pool4"],
    "D": ["drop"],
}
_CONV_KIND = {3: 0, 1: 1, 5: 5, 11: 11}  # kernel size -> maua_net_desc::conv_kind


def _layer_names(channels):
    """models.py:140-243 (vgg16_dict / vgg19_dict): conv{b}_{i}, relu{b}_{i}, pool{b}."""
    names = {"C": [], "R": [], "P": []}
    block, idx = 1, 1
    for c in channels:
        if c == "P":
            names["P"].append(f"pool{block}")
            block, idx = block + 1, 1
        else:
            names["C"].append(f"conv{block}_{idx}")
            names["R"].append(f"relu{block}_{idx}")
            idx += 1
    return names


vgg16_dict = _layer_names(channel_list["VGG-16"])
vgg19_dict = _layer_names(channel_list["VGG-19"])

_MODES = {"none": _lib.MODE_NONE, "capture": _lib.MODE_CAPTURE, "loss": _lib.MODE_LOSS}


def default_impl() -> int:
    """MAUA_PRECISION=tf32 (default): tcgen05 kernels, TF32 operands / FP32 accumulate -- the arithmetic the reference's own
    GPU path uses (cuDNN allows TF32 by default).  MAUA_PRECISION=fp32: the exact-arithmetic kernels (csrc/conv_fp32.cu),
    which reproduce the reference's CPU / fp32 results to ~1e-6 at ~1/30 of the speed."""
    prec = os.environ.get("MAUA_PRECISION", "tf32").lower()
    if prec in ("tf32", ""):
        return _lib.MAUA_IMPL_TC
    if prec in ("fp32", "exact"):
        return _lib.MAUA_IMPL_FP32
    raise ValueError(f"MAUA_PRECISION={prec!r}: expected tf32 or fp32")


def _match_architecture(name: str):
    if "prun" in name:  # models.py:249-258: channel list "VGG-16p" (24, 22, 41, 51, 108, ...)
        return "VGG-16p"
    if "vgg19" in name or "vgg-19" in name:
        return "VGG-19"
    # models.py:259-288, :334-347: the fcn32s / nyud / sod checkpoints are VGG-16 feature stacks with other heads
    if any(k in name for k in ("vgg16", "vgg-16", "fcn32s", "nyud", "sod")):
        return "VGG-16"
    if "vgg" not in name and "nin" in name:  # models.py:327-339
        return "NIN"
    return None


def _architecture(model_file: str, pooling: str):
    """models.py:246-347 for the architectures this backend accelerates: (channels, layer names) from the file name (the
    reference matches substrings of the whole path; the file name is tried first here so that a directory called
    e.g. "episode" does not select the SOD model)."""
    full = str(model_file).lower()
    arch = _match_architecture(os.path.basename(full)) or _match_architecture(full)
    if arch is None:
        raise ValueError("Model architecture not recognized.")  # models.py:341 (VGG-16 / -19 / pruned / NIN are what it knows)
    if pooling not in ("max", "avg"):
        raise ValueError("Unrecognized pooling argseter")  # models.py:124
    if arch == "NIN":
        return NIN_LAYERS, nin_dict
    channels, layer_list = channel_list[arch], (vgg19_dict if arch == "VGG-19" else vgg16_dict)
    return channels, layer_list


# models.py:246-347: a --model_file that is not an existing path is a model NAME; the reference maps it to its model zoo
# (and downloads the file when it is missing -- there is no network on a B200 box, so that part is an error message)
_MODELZOO = [("prun", "vgg16-prune.pth"), ("nyud", "nyud-fcn32s-color-heavy.pth"), ("fcn32s", "fcn32s-heavy-pascal.pth"),
             ("sod", "vgg16-sod.pth"), ("vgg19", "vgg19.pth"), ("vgg16", "vgg16.pth"), ("nin", "nin.pth")]


def resolve_model_file(model_file: str) -> str:
    """The checkpoint path `select_model` would load (models.py:246-347): `model_file` itself when it exists, otherwise
    `modelzoo/<name>.pth` for the model the name selects (relative to the working directory, like the reference; a
    directory can be given with $MAUA_MODELZOO).  The stock config/scaling-img.json names models this way ("vgg19",
    "prune", "nin") and optim.set_model_args overwrites args.model_file with those names for every scale."""
    mf = str(model_file)
    if os.path.exists(mf):
        return mf
    low = mf.lower()
    for key, fname in _MODELZOO:
        if key in low:
            zoo = os.path.join(os.environ.get("MAUA_MODELZOO", "modelzoo"), fname)
            if os.path.exists(zoo):
                return zoo
            raise FileNotFoundError(
                f"model_file {mf!r} is not a file and the model-zoo checkpoint {zoo!r} does not exist either; the reference "
                f"would download it here (models.py:246-347), which is not possible offline: place the checkpoint there, set "
                f"$MAUA_MODELZOO, or pass --model_file with a real path (and name it in the --scaling_args JSON)")
    raise ValueError("Model architecture not recognized.")  # models.py:338


def select_model(model_file: str, pooling: str, verbose: bool, disable_check: bool):
    """Returns (channels, layer names, state dict) -- the checkpoint is read with torch.load like models.py:343."""
    channels, layer_list = _architecture(model_file, pooling)
    sd = torch.load(resolve_model_file(model_file), map_location="cpu")
    return channels, layer_list, sd


def padded_channels(c: int, style_tap: bool) -> int:
    """Channel count a conv layer of `c` real channels runs with: the tcgen05 conv and Gram kernels tile channels in
    multiples of 64 (csrc/plan.cu).  The channel-pruned VGG-16 (models.py:136: 24, 22,
    41, 51, 108, 89, 111, 184, 276, 228, ...) is run with zero weights / bias in the padded channels -- they stay exactly
    zero through ReLU, pooling, Gram and every gradient -- and the plan normalises with the real count
    (maua_net_desc::norm_channels)."""
    return -(-c // 64) * 64


def _pad_params(params, real, padded):
    """Zero-pad [(w [c, cin, 3, 3], b [c])] to the padded channel counts (both the output and the input side)."""
    out, cin_p = [], 3
    for (w, b), c, cp in zip(params, real, padded):
        if (cp, cin_p) != (w.shape[0], w.shape[1]):
            wp = torch.zeros(cp, cin_p, *w.shape[2:], dtype=w.dtype)
            wp[: w.shape[0], : w.shape[1]] = w
            bp = torch.zeros(cp, dtype=b.dtype)
            bp[: b.shape[0]] = b
            w, b = wp, bp
        out.append((w, b))
        cin_p = cp
    return out


def _conv_params(sd, channels, disable_check):
    """(weight, bias) per conv, from torchvision-style keys features.{k}.weight (k counts conv/relu/pool slots)."""
    params, k, cin, conv_i = [], 0, 3, 0
    nin = channels is NIN_LAYERS
    for c in channels:
        if c == "P":
            k += 1
            continue
        c, ks = c if isinstance(c, tuple) else (c, 3)
        if nin:
            k = NIN_FEATURE_INDEX[conv_i]
        wk, bk = f"features.{k}.weight", f"features.{k}.bias"
        if wk not in sd or bk not in sd:
            raise KeyError(f"checkpoint is missing {wk} / {bk}")
        w, b = sd[wk].float(), sd[bk].float()
        if tuple(w.shape) != (c, cin, ks, ks):
            raise ValueError(f"{wk} has shape {tuple(w.shape)}, expected {(c, cin, ks, ks)}")
        params.append((w, b))
        cin = c
        k += 2
        conv_i += 1
    return params


class _PlanFunction(torch.autograd.Function):
    """net(pastiche) as one autograd node: forward -> vector of module losses, backward -> pastiche.grad."""

    @staticmethod
    def forward(ctx, x, net):
        ctx.net = net
        ctx.token = net._forward_plan(x, keep=True)
        return net._loss_vec.clone()

    @staticmethod
    def backward(ctx, grad_losses):
        net = ctx.net
        if ctx.token != net._fwd_token:
            raise RuntimeError("maua_style_b200: backward() called after another forward pass of the same network")
        return net._backward_plan(grad_losses), None


class _PlanCore:
    """What survives from one `load_model` call to the next: the frozen weights on the device(s) and the plan(s) built
    from them (GEMM-layout weight copies, tensor maps, arena).  The reference reloads the checkpoint and rebuilds the
    network for every scale and every video pass (optim.py:128-129 -> models.py:351-453: torch.load + deepcopy + .cuda());
    SURVEY.md section 8f rank 1 asks for this to be cached.  A core is keyed by everything the plan depends on and is
    only handed to one live network at a time."""

    def __init__(self, entries, params, avg_pool, tap_sig, device, bounds, devs, norm_channels=None, conv_kinds=None,
                 pool_kind=0):
        lib = _lib.load()
        self._lib = lib
        self.device = device
        # kept so that sibling cores can be built for the frames of an img_vid window (window.py)
        self.entries, self.avg_pool, self.tap_sig, self.bounds, self.devs = list(entries), avg_pool, list(tap_sig), list(bounds), list(devs)
        self.norm_channels = list(norm_channels) if norm_channels is not None else None
        self.conv_kinds, self.pool_kind = (list(conv_kinds) if conv_kinds is not None else None), int(pool_kind)
        self.weights = nn.ParameterList([nn.Parameter(w.to(device).contiguous(), requires_grad=False) for w, _ in params])
        self.biases = nn.ParameterList([nn.Parameter(b.to(device).contiguous(), requires_grad=False) for _, b in params])
        desc = _lib.NetDesc()
        desc.n_entries = len(entries)
        ci = 0
        for i, c in enumerate(entries):
            desc.channels[i] = c
            if norm_channels is not None:
                desc.norm_channels[i] = norm_channels[i]
            if conv_kinds is not None:
                desc.conv_kind[i] = conv_kinds[i]
            if c > 0:
                desc.weights[i] = self.weights[ci].data_ptr()
                desc.biases[i] = self.biases[ci].data_ptr()
                ci += 1
        desc.avg_pool = int(avg_pool)
        desc.pool_kind = int(pool_kind)
        desc.n_taps = len(tap_sig)
        for t, (ridx, kind) in enumerate(tap_sig):
            desc.tap_relu_index[t] = ridx
            desc.tap_kind[t] = kind
        n_slots = len(tap_sig) + 2
        self.stages = []
        self.loop_states = collections.OrderedDict()  # optim.optimize_device: optimizer state + captured graph per image size
        self.owner = None  # weakref to the network currently using this core
        torch.cuda.synchronize(device)  # the weight uploads above are consumed by plan creation on other devices
        for k, dv in enumerate(devs):
            plan = C.c_void_p()
            with torch.cuda.device(dv):
                if len(devs) == 1:
                    _lib.check(lib.maua_plan_create(dv.index or 0, C.byref(desc), C.byref(plan)), "maua_plan_create")
                else:
                    if k > 0:
                        _lib.check(lib.maua_enable_peer_access(devs[k - 1].index, dv.index), "maua_enable_peer_access")
                    _lib.check(lib.maua_plan_create_stage(dv.index or 0, C.byref(desc), bounds[k], bounds[k + 1],
                                                          C.byref(plan)), "maua_plan_create_stage")
            self.stages.append({"plan": plan, "device": dv, "begin": bounds[k], "end": bounds[k + 1],
                                "loss_vec": torch.zeros(n_slots, device=dv),
                                "coefs": torch.zeros(n_slots, device=dv), "x_in": None, "g_top": None})

    def in_use(self) -> bool:
        return self.owner is not None and self.owner() is not None

    def __del__(self):
        stages, self.stages = getattr(self, "stages", []), []
        for st in stages:
            try:
                self._lib.maua_plan_destroy(st["plan"])
            except Exception:
                pass


_CORE_CACHE: "collections.OrderedDict" = collections.OrderedDict()
_CORE_CACHE_MAX = 2
cache_stats = {"hits": 0, "misses": 0}


def clear_model_cache() -> None:
    """Drop the cached weights / plans (frees their device memory once no network uses them)."""
    _CORE_CACHE.clear()



class B200Net(nn.Module):
    """The truncated feature stack with its loss modules, executed by libmaua_b200 (csrc/plan.cu)."""

    def __init__(self, entries: List[int], params, avg_pool: bool, taps, tv_mod, temporal_mod, device: torch.device,
                 stage_bounds: Optional[List[int]] = None, devices: Optional[List[torch.device]] = None,
                 core: Optional[_PlanCore] = None, norm_channels: Optional[List[int]] = None,
                 conv_kinds: Optional[List[int]] = None, pool_kind: int = 0):
        super().__init__()
        _lib.require_gpu()
        self._lib = _lib.load()
        self.device = device  # device of the image side (stage 0); losses and the image gradient are delivered there
        self.entries = entries
        self.conv_kinds = list(conv_kinds) if conv_kinds is not None else [0] * len(entries)
        self.pool_kind = int(pool_kind)
        self.taps = taps  # [(relu_index, module)] ordered by relu index
        self.tv_mod = tv_mod
        self.temporal_mod = temporal_mod
        self._n_slots = len(taps) + 2
        # stages of the layer-wise split (models.py:503-566); the common case is ONE stage = the whole stack
        bounds = list(stage_bounds) if stage_bounds else [0, len(entries)]
        devs = list(devices) if devices else [device]
        if len(devs) != len(bounds) - 1 or bounds[0] != 0 or bounds[-1] != len(entries) or sorted(set(bounds)) != bounds:
            raise ValueError(f"bad stage layout: bounds {bounds} for {len(devs)} device(s) and {len(entries)} entries")
        if core is None:
            tap_sig = [(ridx, _lib.TAP_STYLE if isinstance(mod, StyleLoss) else _lib.TAP_CONTENT) for ridx, mod in taps]
            core = _PlanCore(entries, params, avg_pool, tap_sig, device, bounds, devs, norm_channels, conv_kinds, pool_kind)
        core.owner = weakref.ref(self)
        self._core = core
        # parameters are kept (frozen) so that net.parameters() / state inspection behave like the reference's net
        self.weights, self.biases = core.weights, core.biases
        self._stages = core.stages
        for st in self._stages:  # a re-used core starts from the library defaults
            st["loss_vec"].zero_()
            _lib.check(self._lib.maua_plan_set_impl(st["plan"], default_impl()), "maua_plan_set_impl")
            _lib.check(self._lib.maua_plan_set_profile(st["plan"], 0), "maua_plan_set_profile")
            _lib.check(self._lib.maua_plan_set_conv_tail(st["plan"], 1 if os.environ.get("MAUA_SPLITK", "0") == "1"
                                                         else int(os.environ.get("MAUA_CONV_TAIL", "2"))), "maua_plan_set_conv_tail")
            _lib.check(self._lib.maua_plan_set_fuse_pool(st["plan"], int(os.environ.get("MAUA_FUSE_POOL", "1") != "0")),
                       "maua_plan_set_fuse_pool")
        self._plan = self._stages[0]["plan"]
        self._impl = default_impl()
        self._window = None       # FrameWindow (window.py): evaluation of [B > 1, 3, H, W] inputs (img_vid)
        self._window_live = False
        self._loss_vec = torch.zeros(self._n_slots, device=device)
        self._coefs = self._stages[0]["coefs"]
        self._fwd_token = 0
        self.reuse_target_buffers = False  # set by optimisation loops that keep a captured graph across images
        self._tap_channels = []   # real channel count of each tapped layer (what .target exposes)
        self._tap_cpad = []       # channel count the plan runs it with (>= real: zero-padded, see padded_channels)
        self._tap_stage = []
        ch = [c for c in entries if c > 0]
        real = [c for c in (norm_channels or entries) if c > 0]
        self.real_channels = real
        conv_entry = [i for i, c in enumerate(entries) if c > 0]
        for ridx, _ in taps:
            self._tap_channels.append(real[ridx])
            self._tap_cpad.append(ch[ridx])
            e = conv_entry[ridx]
            self._tap_stage.append(next(k for k, st in enumerate(self._stages) if st["begin"] <= e < st["end"]))
        self.content_losses, self.style_losses, self.tv_losses, self.temporal_losses = [], [], [], []

    def __del__(self):
        # the plans belong to the core (shared with the model cache); dropping the last reference destroys them
        try:
            self.__dict__.update(_stages=[], _plan=None, _core=None)
        except Exception:  # interpreter shutdown
            pass

    @property
    def n_stages(self) -> int:
        return len(self._stages)

    def _tap_device(self, t: int) -> torch.device:
        return self._stages[self._tap_stage[t]]["device"]

    # ------------------------------------------------------------------------------------------------
    def set_impl(self, impl: int):
        self._impl = impl
        for st in self._stages:
            _lib.check(self._lib.maua_plan_set_impl(st["plan"], impl), "maua_plan_set_impl")

    def set_fuse_pool(self, enable: bool):
        """Pool inside the producing conv's epilogue (csrc/conv_tc.cu, the default) or as a separate pass."""
        for st in self._stages:
            _lib.check(self._lib.maua_plan_set_fuse_pool(st["plan"], int(enable)), "maua_plan_set_fuse_pool")

    def set_side_stream(self, mode: int):
        """Loss modules of the forward pass (Gram + StyleLoss, ContentLoss MSE) on a side stream next to the following
        convolutions: -1 automatic (on up to 640 x 640 pixels), 0 off, 1 on.  Same kernels, bit-identical results."""
        for st in self._stages:
            _lib.check(self._lib.maua_plan_set_side_stream(st["plan"], int(mode)), "maua_plan_set_side_stream")

    def set_splitk(self, enable: bool):
        """K-split of the last partial wave of conv tiles (csrc/conv_tc.cu; off by default)."""
        for st in self._stages:
            _lib.check(self._lib.maua_plan_set_splitk(st["plan"], int(enable)), "maua_plan_set_splitk")

    def set_conv_tail(self, mode: int):
        """Last partial wave of the persistent conv kernels: 0 whole tiles, 1 K-split, 2 half-N items (default)."""
        for st in self._stages:
            _lib.check(self._lib.maua_plan_set_conv_tail(st["plan"], int(mode)), "maua_plan_set_conv_tail")

    def device_bytes(self) -> int:
        return sum(int(self._lib.maua_plan_device_bytes(st["plan"])) for st in self._stages)

    def last_launches(self):
        tf = tb = 0
        for st in self._stages:
            f, b = C.c_int(), C.c_int()
            _lib.check(self._lib.maua_plan_last_launches(st["plan"], C.byref(f), C.byref(b)))
            tf, tb = tf + f.value, tb + b.value
        return tf, tb

    def set_profile(self, enable: bool):
        _lib.check(self._lib.maua_plan_set_profile(self._plan, int(enable)), "maua_plan_set_profile")

    def profile(self):
        """Per-launch records of the last forward + backward (needs set_profile(True)); synchronises the stream."""
        import json

        cap = 1 << 16
        buf = C.create_string_buffer(cap)
        n = self._lib.maua_plan_profile_json(self._plan, buf, C.c_long(cap), _lib.stream_ptr())
        if n < 0 or n > cap:
            raise RuntimeError("maua_plan_profile_json failed")
        return json.loads(buf.value.decode())

    def slot_modules(self):
        """Loss modules in plan slot order: taps..., TV, temporal (None where absent)."""
        return [m for _, m in self.taps] + [self.tv_mod, self.temporal_mod]

    def _tap_hw(self, H, W, ridx):
        h, w, conv = H, W, -1
        for c, kind in zip(self.entries, self.conv_kinds):
            if c == 0:
                if self.pool_kind == 1:  # 3x3 / stride 2 / ceil_mode (models.py:77-80)
                    h, w = (0 if h < 2 else (h - 2) // 2 + 1), (0 if w < 2 else (w - 2) // 2 + 1)
                else:
                    h, w = h // 2, w // 2
            else:
                if kind == 11:   # 11x11 / stride 4 / no padding (models.py:83)
                    h, w = (h - 11) // 4 + 1, (w - 11) // 4 + 1
                conv += 1
                if conv == ridx:
                    return h, w
        raise AssertionError

    def _build_io(self, H: int, W: int, window: bool = False):
        """The per-forward description of every loss module for the plan (maua_tap_io / maua_image_io): modes, strengths,
        target buffers.  Capture-mode modules get their target buffers here (loss.py:61-62, :146-151)."""
        tio = (_lib.TapIO * max(len(self.taps), 1))()
        for t, (ridx, mod) in enumerate(self.taps):
            io = tio[t]
            io.mode = _MODES[mod.mode]
            C_ = self._tap_channels[t]
            Cp = self._tap_cpad[t]
            tdev = self._tap_device(t)
            if isinstance(mod, StyleLoss):
                io.use_covariance = int(bool(mod.use_covariance))
                vsf = self._effective_vsf(mod, C_)
                io.value_scale = float(mod.strength) * (1.0 + vsf)
                if mod.mode == "capture" and not window:  # (a window of B > 1 frames manages its captures itself)
                    fresh = mod.target.nelement() == 0
                    if fresh and Cp == C_:
                        if self.reuse_target_buffers:
                            # optimisation loops (optim.optimize_device) build a new network per call on the SAME plan core: the
                            # style targets go into the core's buffers again, so that the iteration captured for the previous
                            # call launches identical kernel arguments and its CUDA graph is replayed instead of re-captured
                            bufs = self._core.__dict__.setdefault("style_target_bufs", {})
                            if (t, C_) not in bufs:
                                bufs[(t, C_)] = torch.zeros(C_, C_, device=tdev)
                            mod.target = bufs[(t, C_)]
                        else:
                            mod.target = torch.zeros(C_, C_, device=tdev)
                    io.capture_weight = float(mod.blend_weight)
                    io.capture_accumulate = 0 if fresh else 1
                if mod.mode != "none" and Cp != C_:
                    buf = self._padded_style_target(mod, C_, Cp, tdev)  # mod.target is its [:C, :C] view
                    io.target = buf.data_ptr()
                    io.target_elems = buf.numel()
                elif mod.mode != "none":
                    if not (mod.target.device == tdev and mod.target.dtype == torch.float32 and mod.target.is_contiguous()):
                        mod.target = mod.target.to(tdev, torch.float32).contiguous()
                    io.target = mod.target.data_ptr()
                    io.target_elems = mod.target.numel()
                if mod.mode == "capture" and not window:
                    if not self._dynamic_skipped(mod, C_):
                        mod.video_target = mod.target  # identical for B = 1 (loss.py:164-175)
                    mod.loss = 0
            elif Cp != C_:
                io.value_scale = float(mod.strength)
                h, w = self._tap_hw(H, W, ridx)
                buf = self._padded_content_target(mod, C_, Cp, h, w, tdev, capture=mod.mode == "capture")
                if mod.mode != "none" and buf is not None:
                    io.target = buf.data_ptr()
                    io.target_elems = buf.numel()
            else:
                io.value_scale = float(mod.strength)
                h, w = self._tap_hw(H, W, ridx)
                if mod.mode == "capture":
                    # NCHW-shaped view over NHWC memory (channels_last): .size() matches the reference's target.  With
                    # reuse_target_buffers (optimisation loops over many images of one size) an existing buffer of the right
                    # shape is overwritten in place, so that a captured CUDA graph of the iteration stays valid.
                    tgt = mod.target
                    if not (self.reuse_target_buffers and tgt.nelement() != 0 and tuple(tgt.shape) == (1, C_, h, w)
                            and tgt.device == tdev and tgt.permute(0, 2, 3, 1).is_contiguous()):
                        mod.target = torch.empty(1, h, w, C_, device=tdev).permute(0, 3, 1, 2)
                if mod.mode != "none" and mod.target.nelement() != 0:
                    tgt = mod.target
                    if tuple(tgt.shape[1:]) == (C_, h, w):
                        if not (tgt.device == tdev and tgt.permute(0, 2, 3, 1).is_contiguous()):
                            tgt = tgt.to(tdev, torch.float32).contiguous(memory_format=torch.channels_last)
                            mod.target = tgt
                        io.target = tgt.data_ptr()
                        io.target_elems = tgt.numel()
                    # else: shape mismatch -> module is skipped, as loss.py:44 does
        iio = _lib.ImageIO()
        if self.tv_mod is not None:
            iio.tv_mode = _lib.MODE_LOSS  # TVLoss.loss is assigned on every forward (loss.py:232)
            iio.tv_strength = float(self.tv_mod.strength)
        tm = self.temporal_mod
        if tm is not None and tm.mode != "none":
            iio.temporal_mode = _MODES[tm.mode]
            iio.temporal_strength = float(tm.strength)
            if tm.mode == "capture":
                if not (self.reuse_target_buffers and tm.target.nelement() != 0 and tuple(tm.target.shape) == (1, 3, H, W)
                        and tm.target.is_cuda and tm.target.is_contiguous()):
                    tm.target = torch.empty(1, 3, H, W, device=self.device)
            if tm.target.nelement() != 0 and tuple(tm.target.shape[1:]) == (3, H, W):
                if not (tm.target.is_cuda and tm.target.is_contiguous()):
                    tm.target = tm.target.to(self.device, torch.float32).contiguous()
                iio.temporal_target = tm.target.data_ptr()
                iio.temporal_target_elems = tm.target.numel()
                if tm.weights is not None:
                    wts = tm.weights
                    if wts.numel() != H * W:
                        raise ValueError("temporal weights must have H*W elements ([1,1,H,W])")
                    if not (wts.is_cuda and wts.dtype == torch.float32 and wts.is_contiguous()):
                        wts = wts.to(self.device, torch.float32).contiguous()
                        tm.weights = wts
                    iio.temporal_weights = wts.data_ptr()
        return tio, iio

    # -- zero-padded layers (channel-pruned VGG-16): the plan's targets are padded buffers, `.target` is the real-size view --
    @staticmethod
    def _padded_style_target(mod, C_, Cp, tdev):
        buf = getattr(mod, "_b200_target_pad", None)
        tgt = mod.target
        aliased = (buf is not None and tuple(buf.shape) == (Cp, Cp) and buf.device == tdev and tgt.nelement() != 0
                   and tgt.data_ptr() == buf.data_ptr() and tuple(tgt.shape) == (C_, C_) and tgt.stride() == (Cp, 1))
        if not aliased:
            buf = torch.zeros(Cp, Cp, device=tdev)
            if tgt.nelement() != 0:  # a target set from outside (or moved): adopt its values
                buf[:C_, :C_] = tgt.to(tdev, torch.float32)
            mod._b200_target_pad = buf
            mod.target = buf[:C_, :C_]
        return buf

    def _padded_content_target(self, mod, C_, Cp, h, w, tdev, capture):
        buf = getattr(mod, "_b200_target_pad", None)
        tgt = mod.target
        aliased = (buf is not None and tuple(buf.shape) == (1, h, w, Cp) and buf.device == tdev and tgt.nelement() != 0
                   and tgt.data_ptr() == buf.data_ptr() and tuple(tgt.shape) == (1, C_, h, w))
        if aliased and (not capture or self.reuse_target_buffers):
            return buf
        if capture:
            buf = torch.zeros(1, h, w, Cp, device=tdev)
        elif tgt.nelement() != 0 and tuple(tgt.shape[1:]) == (C_, h, w):
            buf = torch.zeros(1, h, w, Cp, device=tdev)
            buf[..., :C_] = tgt.to(tdev, torch.float32).permute(0, 2, 3, 1)
        else:
            return None  # no target yet, or a shape mismatch: the module is skipped (loss.py:44)
        mod._b200_target_pad = buf
        mod.target = buf.permute(0, 3, 1, 2)[:, :C_]
        return buf

    @staticmethod
    def _dynamic_skipped(mod, C_) -> bool:
        """loss.py:165-166: a video target captured from windows of B > 1 frames is [B*C, B*C]; a single image then has no
        dynamic term (neither captured nor evaluated)."""
        vt = mod.video_target
        return vt.nelement() != 0 and vt.shape[0] != C_

    def _effective_vsf(self, mod, C_) -> float:
        vsf = float(mod.video_style_factor)
        return vsf if vsf > 0 and not self._dynamic_skipped(mod, C_) else 0.0

    def io_signature(self, H: int, W: int) -> bytes:
        """Everything a captured iteration bakes into its kernel arguments besides the pastiche: extents, module modes,
        strengths and target addresses.  Two iterations with equal signatures launch identical kernels."""
        tio, iio = self._build_io(H, W)
        # (the plan's workspaces are part of the captured kernel arguments too: they are re-allocated when a larger image comes)
        gen = ",".join(str(int(self._lib.maua_plan_workspace_generation(st["plan"]))) for st in self._stages)
        return bytes(memoryview(tio)) + bytes(memoryview(iio)) + f"{H}x{W} ws {gen} impl {self._impl}".encode()

    def _forward_plan(self, x: torch.Tensor, keep: bool) -> int:
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected a [B,3,H,W] image batch, got {tuple(x.shape)}")
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise ValueError("internal: image must be a contiguous fp32 CUDA tensor")
        self._window_live = x.shape[0] != 1
        if self._window_live:
            # a window of B > 1 frames (img_vid: loss.py:141-181 batch semantics): one plan per frame + one SYRK over all
            if self._window is None:
                from .window import FrameWindow

                self._window = FrameWindow(self)
            self._keepalive = (x, None, None)
            self._window.forward(x, keep)
            self._fwd_token += 1
            return self._fwd_token
        H, W = int(x.shape[2]), int(x.shape[3])
        tio, iio = self._build_io(H, W)
        self._keepalive = (x, tio, iio)
        if len(self._stages) == 1:
            with torch.cuda.device(self.device):
                _lib.check(self._lib.maua_plan_forward(self._plan, _lib.ptr(x), H, W, tio, C.byref(iio), _lib.ptr(self._loss_vec),
                                                       int(keep), _lib.stream_ptr()), "maua_plan_forward")
        else:
            self._forward_stages(x, H, W, tio, iio, keep)
        self._fwd_token += 1
        return self._fwd_token

    def _forward_stages(self, x, H, W, tio, iio, keep):
        """models.py:517-525 ModelParallel.forward: the stages run one after the other, each on its own device and stream.
        The hand-over tensor lives on the CONSUMING device and is written there directly by the producing stage's last
        kernel (peer stores over NVLink); an event orders the two streams, like the reference's `.to(device)`."""
        last_live = max([self._tap_stage[t] for t, (_, m) in enumerate(self.taps) if m.mode != "none"], default=0)
        self._last_live_stage = last_live
        inp, h, w = x, H, W
        prev_event = None
        for k, st in enumerate(self._stages[:last_live + 1]):
            dv = st["device"]
            with torch.cuda.device(dv):
                stream = torch.cuda.current_stream(dv)
                if prev_event is not None:
                    stream.wait_event(prev_event)
                boundary = None
                if k < last_live:
                    oh, ow, oc = C.c_int(), C.c_int(), C.c_int()
                    _lib.check(self._lib.maua_plan_stage_output_shape(st["plan"], h, w, C.byref(oh), C.byref(ow), C.byref(oc)))
                    nxt = self._stages[k + 1]
                    shape = (oh.value, ow.value, oc.value)
                    if nxt["x_in"] is None or tuple(nxt["x_in"].shape) != shape:
                        nxt["x_in"] = torch.empty(shape, device=nxt["device"])          # input of stage k+1, on ITS device
                        st["g_top"] = torch.empty(shape, device=dv)                    # its gradient, on THIS device
                    boundary = nxt["x_in"]
                _lib.check(self._lib.maua_plan_forward_stage(st["plan"], _lib.ptr(inp), h, w, tio, C.byref(iio),
                                                             _lib.ptr(st["loss_vec"]), int(keep), _lib.ptr(boundary),
                                                             C.c_void_p(stream.cuda_stream)), "maua_plan_forward_stage")
                prev_event = torch.cuda.Event()
                prev_event.record(stream)
            if boundary is not None:
                inp, h, w = boundary, boundary.shape[0], boundary.shape[1]
        # the module losses are collected on the image-side device (optim.py:211 `mod.loss.to(backward_device)`)
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            main.wait_event(prev_event)
            total = self._stages[0]["loss_vec"].clone()
            for st in self._stages[1:last_live + 1]:
                total += st["loss_vec"].to(self.device, non_blocking=True)
            self._loss_vec.copy_(total)

    def _backward_plan(self, grad_losses: torch.Tensor) -> torch.Tensor:
        x = self._keepalive[0]
        if self._window_live:
            return self._window.backward(grad_losses, x)
        n = self._n_slots
        strength = (C.c_float * n)()
        vsf = (C.c_float * n)()
        normalize = (C.c_int * n)()
        kind = (C.c_int * n)()
        for i, mod in enumerate(self.slot_modules()):
            if mod is None:
                continue
            strength[i] = float(mod.strength)
            if isinstance(mod, StyleLoss):
                kind[i], vsf[i], normalize[i] = 0, self._effective_vsf(mod, self._tap_channels[i]), int(bool(mod.normalize))
            elif isinstance(mod, ContentLoss):
                kind[i], normalize[i] = 1, int(bool(mod.normalize))
            else:
                kind[i] = 2
        up = grad_losses.detach().to(self.device, torch.float32).contiguous()
        grad = torch.empty_like(x)
        if len(self._stages) == 1:
            with torch.cuda.device(self.device):
                _lib.check(self._lib.maua_loss_grad_coefs(_lib.ptr(up), _lib.ptr(self._coefs), n, strength, vsf, normalize, kind,
                                                          _lib.stream_ptr()), "maua_loss_grad_coefs")
                _lib.check(self._lib.maua_plan_backward(self._plan, _lib.ptr(self._coefs), _lib.ptr(grad), _lib.stream_ptr()),
                           "maua_plan_backward")
            return grad
        # layer-wise split: last live stage first; each stage stores d/d(its input) straight into the previous stage's
        # g_top buffer (peer memory) and an event hands the stream order over
        last_live = self._last_live_stage
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)  # `up` (and the forward) are ordered on the image-side stream
        prev_event = ready
        for k in range(last_live, -1, -1):
            st = self._stages[k]
            dv = st["device"]
            with torch.cuda.device(dv):
                stream = torch.cuda.current_stream(dv)
                stream.wait_event(prev_event)
                up_k = up if dv == self.device else up.to(dv, non_blocking=True)
                _lib.check(self._lib.maua_loss_grad_coefs(_lib.ptr(up_k), _lib.ptr(st["coefs"]), n, strength, vsf, normalize, kind,
                                                          C.c_void_p(stream.cuda_stream)), "maua_loss_grad_coefs")
                g_top = st["g_top"] if k < last_live else None
                out = grad if k == 0 else self._stages[k - 1]["g_top"]
                _lib.check(self._lib.maua_plan_backward_stage(st["plan"], _lib.ptr(st["coefs"]), _lib.ptr(g_top), _lib.ptr(out),
                                                              C.c_void_p(stream.cuda_stream)), "maua_plan_backward_stage")
                st["_keep"] = up_k
                prev_event = torch.cuda.Event()
                prev_event.record(stream)
        return grad

    def _live_slots(self):
        """Slots whose module produced a loss in the last forward (mode 'loss' and not skipped)."""
        live = []
        for i, mod in enumerate(self.slot_modules()):
            if mod is None:
                continue
            if isinstance(mod, TVLoss):
                live.append(i)
            elif mod.mode == "loss" and mod.target.nelement() != 0:
                live.append(i)
        return live

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        x = input
        if not x.is_cuda or x.device != self.device:
            x = x.to(self.device)
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        any_loss = any(m is not None and getattr(m, "mode", "none") == "loss" for m in self.slot_modules())
        if any_loss and x.requires_grad and torch.is_grad_enabled():
            vec = _PlanFunction.apply(x, self)
        else:
            with torch.no_grad():
                self._forward_plan(x.detach(), keep=False)
            vec = self._loss_vec.clone()
        # hand the module losses out the way the reference modules do: `.loss` accumulates (loss.py:157,181)
        for i in self._live_slots():
            mod = self.slot_modules()[i]
            if isinstance(mod, TVLoss):
                mod.loss = vec[i]  # assigned, not accumulated (loss.py:232)
            elif isinstance(mod, ContentLoss):
                mod.loss = vec[i]  # loss.py:49 resets to 0 before accumulating over the batch
            else:
                mod.loss = mod.loss + vec[i]
        return input

    # -- debugging / tests ---------------------------------------------------------------------------------
    def tap_feature(self, t: int) -> torch.Tensor:
        """Feature map of tap t from the last forward as an NCHW tensor (copy)."""
        h, w, c = C.c_int(), C.c_int(), C.c_int()
        null = C.c_void_p(0)
        st = self._stages[self._tap_stage[t]]
        with torch.cuda.device(st["device"]):
            _lib.check(self._lib.maua_plan_tap_feature(st["plan"], t, null, C.byref(h), C.byref(w), C.byref(c), _lib.stream_ptr()))
            out = torch.empty(1, h.value, w.value, c.value, device=st["device"])
            _lib.check(self._lib.maua_plan_tap_feature(st["plan"], t, _lib.ptr(out), C.byref(h), C.byref(w), C.byref(c),
                                                       _lib.stream_ptr()))
        return out.permute(0, 3, 1, 2)[:, : self._tap_channels[t]].contiguous()

    def entry_output(self, i: int) -> torch.Tensor:
        """Output of stack entry i (conv: post-ReLU activation, pool: pooled map) from the last forward, NCHW (copy)."""
        k = next(k for k, st in enumerate(self._stages) if st["begin"] <= i < st["end"])
        st = self._stages[k]
        h, w, c, ip = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        with torch.cuda.device(st["device"]):
            _lib.check(self._lib.maua_plan_entry_output(st["plan"], i - st["begin"], C.c_void_p(0), C.byref(h), C.byref(w),
                                                        C.byref(c), C.byref(ip), _lib.stream_ptr()))
            out = torch.empty(1, h.value, w.value, c.value, device=st["device"])
            _lib.check(self._lib.maua_plan_entry_output(st["plan"], i - st["begin"], _lib.ptr(out), C.byref(h), C.byref(w),
                                                        C.byref(c), C.byref(ip), _lib.stream_ptr()))
        return out.permute(0, 3, 1, 2).contiguous()

    def tap_gram(self, t: int) -> torch.Tensor:
        """Normalised Gram / covariance matrix of style tap t from the last forward (copy)."""
        c = C.c_int()
        st = self._stages[self._tap_stage[t]]
        with torch.cuda.device(st["device"]):
            _lib.check(self._lib.maua_plan_tap_gram(st["plan"], t, C.c_void_p(0), C.byref(c), _lib.stream_ptr()))
            out = torch.empty(c.value, c.value, device=st["device"])
            _lib.check(self._lib.maua_plan_tap_gram(st["plan"], t, _lib.ptr(out), C.byref(c), _lib.stream_ptr()))
        return out[: self._tap_channels[t], : self._tap_channels[t]].contiguous()


def _device_from_args(args) -> torch.device:
    gpu = str(getattr(args, "gpu", "0")).lower()
    if "c" in gpu.split(",")[0]:
        raise RuntimeError("maua_style_b200 has no CPU path: --gpu c is not supported (use the reference's torch modules)")
    return torch.device("cuda", int(gpu.split(",")[0]))


def build_net(args, entries, params_fn, taps, tv_mod, temporal_mod, device, stage_bounds=None, devices=None,
              model_path=None, norm_channels=None, conv_kinds=None, pool_kind=0) -> "B200Net":
    """A B200Net on a cached plan core when one exists for (checkpoint file, layer layout, taps, pooling, devices) and no
    live network is using it; otherwise on a new core built from `params_fn()` (which reads the checkpoint)."""
    bounds = list(stage_bounds) if stage_bounds else [0, len(entries)]
    devs = list(devices) if devices else [device]
    avg = args.pooling == "avg"
    tap_sig = tuple((ridx, _lib.TAP_STYLE if isinstance(mod, StyleLoss) else _lib.TAP_CONTENT) for ridx, mod in taps)
    key = None
    if os.environ.get("MAUA_NO_MODEL_CACHE", "0") != "1":
        try:
            model_path = model_path or resolve_model_file(str(args.model_file))
            st = os.stat(model_path)
            key = (os.path.realpath(model_path), st.st_mtime_ns, st.st_size, avg, tuple(entries), tuple(norm_channels or ()),
                   tuple(conv_kinds or ()), pool_kind, tap_sig, tuple(bounds), tuple(str(d) for d in devs))
        except (OSError, ValueError):
            key = None
    core = _CORE_CACHE.get(key) if key is not None else None
    if core is not None and core.in_use():
        core = None  # a live network still runs on it: build a fresh core (it takes the cache slot; the old one lives on
        #              with its network)
    if core is None:
        cache_stats["misses"] += 1
        _lib.require_gpu()
        core = _PlanCore(entries, params_fn(), avg, list(tap_sig), device, bounds, devs, norm_channels, conv_kinds, pool_kind)
        if key is not None:
            _CORE_CACHE[key] = core
            while len(_CORE_CACHE) > _CORE_CACHE_MAX:
                _CORE_CACHE.popitem(last=False)
    else:
        cache_stats["hits"] += 1
        _CORE_CACHE.move_to_end(key)
    return B200Net(entries, None, avg, taps, tv_mod, temporal_mod, device, stage_bounds=bounds, devices=devs, core=core,
                   norm_channels=norm_channels, conv_kinds=conv_kinds, pool_kind=pool_kind)


def load_model(args):
    """models.py:351-453."""
    channels, layer_list = _architecture(str(args.model_file), args.pooling)
    model_path = resolve_model_file(str(args.model_file))
    device = _device_from_args(args)
    content_layers = args.content_layers.split(",")
    style_layers = args.style_layers.split(",")
    for nm in content_layers + style_layers:
        if nm.startswith("conv"):
            raise ValueError(f"tap {nm!r}: only relu* taps are supported (conv* taps alias the in-place ReLU buffer in the "
                             "reference, models.py:130)")

    content_losses, style_losses, tv_losses, temporal_losses = [], [], [], []
    n_modules = 0
    tv_mod = temporal_mod = None
    if args.tv_weight > 0:
        tv_mod = TVLoss(args.tv_weight)
        tv_mod.name = f"tv {n_modules}"
        n_modules += 1
        tv_losses.append(tv_mod)
    if args.temporal_weight > 0:
        temporal_mod = ContentLoss(args.temporal_weight, args.normalize_gradients)
        temporal_mod.name = f"temporal {n_modules}"
        n_modules += 1
        temporal_losses.append(temporal_mod)

    entries, taps, conv_kinds = [], [], []
    pool_kind = 1 if channels is NIN_LAYERS else 0
    next_content_idx, next_style_idx, r, conv_i = 1, 1, 0, 0
    for c in channels:
        if not (next_content_idx <= len(content_layers) or next_style_idx <= len(style_layers)):
            break  # models.py:382: nothing below the last requested tap is built
        if c == "P":
            entries.append(0)
            conv_kinds.append(0)
            n_modules += 1
            continue
        c, ks = c if isinstance(c, tuple) else (c, 3)
        entries.append(c)
        conv_kinds.append(_CONV_KIND[ks])
        conv_i += 1
        n_modules += 2  # conv + relu
        name = layer_list["R"][r]
        if name in content_layers:
            if getattr(args, "verbose", False):
                print("Setting up content layer " + str(n_modules) + ": " + name)
            mod = ContentLoss(args.content_weight, args.normalize_gradients)
            mod.name = f"cont {n_modules}"
            n_modules += 1
            content_losses.append(mod)
            taps.append((r, mod))
            next_content_idx += 1
        if name in style_layers:
            if getattr(args, "verbose", False):
                print("Setting up style layer " + str(n_modules) + ": " + name)
            mod = StyleLoss(args.style_weight, args.use_covariance, args.normalize_gradients,
                            video_style_factor=args.video_style_factor, shift_factor=getattr(args, "shift_factor", 0))
            mod.name = f"style {n_modules}"
            n_modules += 1
            style_losses.append(mod)
            taps.append((r, mod))
            next_style_idx += 1
        r += 1
    while entries and entries[-1] == 0:
        entries.pop()  # a trailing pool feeds nothing
        conv_kinds.pop()

    n_convs = conv_i
    # layers whose channel counts the tcgen05 kernels do not tile (channel-pruned VGG-16) run zero-padded
    style_relu = {ridx for ridx, mod in taps if isinstance(mod, StyleLoss)}
    real, padded, k = [c for c in entries if c > 0], [], 0
    for c in real:
        padded.append(padded_channels(c, k in style_relu))
        k += 1
    norm_channels = None
    if padded != real:
        it = iter(padded)
        norm_channels = list(entries)
        entries = [next(it) if c > 0 else 0 for c in entries]

    def params():
        """The checkpoint's conv weights up to the last tapped layer -- only read when no cached core exists."""
        sd = torch.load(model_path, map_location="cpu")  # models.py:343
        raw = _conv_params(sd, channels, getattr(args, "disable_check", False))[:n_convs]
        return _pad_params(raw, real, padded) if norm_channels is not None else raw

    if getattr(args, "multidevice", False):
        from .parallel import setup_multi_device  # layer-wise split over NVLink peers (models.py:537-566)

        return setup_multi_device(entries, params, args, taps, tv_mod, temporal_mod, content_losses, style_losses,
                                  tv_losses, temporal_losses, norm_channels=norm_channels, conv_kinds=conv_kinds, pool_kind=pool_kind)

    net = build_net(args, entries, params, taps, tv_mod, temporal_mod, device, model_path=model_path, norm_channels=norm_channels,
                    conv_kinds=conv_kinds, pool_kind=pool_kind)
    net.content_losses = content_losses
    net.style_losses = style_losses
    net.tv_losses = tv_losses
    net.temporal_losses = temporal_losses
    return net, content_losses + style_losses + tv_losses + temporal_losses
