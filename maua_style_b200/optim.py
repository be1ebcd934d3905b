"""`optimize(content, styles, init, num_iters, args, net=None, losses=None)` with the reference's interface
(reference optim.py:111-255) on top of the fused sm_100a plan.

Kept from the reference: the target-capture protocol (optim.py:22-66), per-size scaling args (optim.py:93-108), the
Adam loop that runs num_iters + 1 evaluate-and-step rounds (`while i[0] <= iters`, optim.py:240) and the single
L-BFGS `step()` of num_iters iterations without line search (optim.py:180-191), `--normalize_weights`
(optim.py:176-178), `--print_iter` / `--save_iter` hooks, and the CPU tensor it returns (optim.py:249).

Different by design: the loop body never synchronises with the host.  feval is two C calls (maua_plan_forward /
maua_plan_backward), the Adam / L-BFGS pixel update is our own kernel (maua_adam_step / maua_lbfgs_step) and the
per-module `.cpu().item()` bookkeeping of optim.py:210 (written to a list nobody reads) only happens when
`--print_iter` asks for a number.  The reference's generic path also still works: `net(pastiche)` is autograd
aware, so `torch.optim.Adam([pastiche])` with the reference's own closure runs unchanged.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import sys
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, models


class _Trace:
    """MAUA_TRACE=1: wall-clock phases of an optimize() call (device synchronised at every mark) on stderr -- where the time of
    a whole multi-resolution job goes besides the iterations themselves."""

    def __init__(self, what: str):
        import os
        import time

        self.on = os.environ.get("MAUA_TRACE", "0") in ("1", "2")
        self.fine = os.environ.get("MAUA_TRACE", "0") == "2"  # also every 100 iterations of the loop
        self.what, self.t, self.time = what, 0.0, time
        if self.on:
            torch.cuda.synchronize()
            self.t = time.perf_counter()

    def mark(self, label: str) -> None:
        if self.on:
            torch.cuda.synchronize()
            now = self.time.perf_counter()
            print(f"[maua trace] {self.what}: {label} {1e3 * (now - self.t):.2f} ms", file=sys.stderr, flush=True)
            self.t = now


def set_content_targets(net, content_image, args=None):
    """optim.py:22-32."""
    for i in net.content_losses:
        i.mode = "capture"
    net(content_image)
    for i in net.content_losses:
        i.mode = "none"


def set_temporal_targets(net, warp_image, warp_weights=None, args=None):
    """optim.py:35-47."""
    for i in net.temporal_losses:
        i.mode = "capture"
        if warp_weights is not None:
            i.weights = warp_weights.to(net.device, torch.float32).contiguous()
    net(warp_image)
    for i in net.temporal_losses:
        i.mode = "none"


def _style_signature(net, style_images, args):
    """Everything the captured style targets depend on: the image tensors themselves (identity, storage, in-place version,
    shape), the capture settings and the arithmetic mode of the plan (`set_impl`)."""
    imgs = tuple((id(t), t.data_ptr(), t._version, tuple(t.shape), str(t.device)) for t in style_images)
    mods = tuple((bool(j.use_covariance), id(j)) for j in net.style_losses)
    # (the arithmetic mode is part of it: targets captured by the TF32 kernels differ from the exact-mode ones by ~1e-4)
    return imgs, tuple(float(w) for w in args.style_blend_weights[:len(style_images)]), mods, getattr(net, "_impl", None)


def set_style_targets(net, style_images, args):
    """optim.py:50-66.  The reference re-captures the style targets on every `optimize` call -- for vid_img that is one
    forward per style image per frame although the style images do not change within a scale (style.py:170-177,
    SURVEY.md section 8f rank 1).  Here a capture is skipped when the same tensors (unmodified) were captured into the
    same modules with the same blend weights; MAUA_NO_TARGET_CACHE=1 restores the unconditional behaviour."""
    import os

    sig = _style_signature(net, style_images, args)
    cached = getattr(net, "_style_cache", None)
    if (cached is not None and cached[0] == sig and os.environ.get("MAUA_NO_TARGET_CACHE", "0") != "1"
            and all(j.target.nelement() != 0 for j in net.style_losses)):
        net.style_cache_hits = getattr(net, "style_cache_hits", 0) + 1
        for j in net.style_losses:
            j.mode = "none"
        return
    for j in net.style_losses:
        j.reset_targets()
        j.mode = "capture"
    for i, image in enumerate(style_images):
        for j in net.style_losses:
            j.blend_weight = args.style_blend_weights[i]
        net(image)
    for j in net.style_losses:
        j.mode = "none"
    # keep the images alive while their signature is cached, so neither id() nor the storage address can be re-used
    net._style_cache = (sig, list(style_images))


def set_style_video_targets(net, style_videos, args):
    """optim.py:69-90 (img_vid): every style video is averaged over all of its windows of `gram_frame_window` frames; a
    forward over a window of B frames captures the per-frame static Grams and the [B*C, B*C] dynamic Gram (window.py)."""
    gfw = int(args.gram_frame_window)
    for j in net.style_losses:
        j.reset_targets()
        j.mode = "capture"
    for i, video in enumerate(style_videos):
        n_win = max(len(video) - gfw + 1, 1)
        for j in net.style_losses:
            j.blend_weight = args.style_blend_weights[i] / n_win
        for window_start in range(n_win):
            net(video[window_start:window_start + gfw])
    for j in net.style_losses:
        j.mode = "none"
    net._style_cache = None  # (the image-style target cache of set_style_targets does not describe these targets)


def wrapping_slice(tensor, start, length, return_indices=False):
    """utils.py:76-85: `length` consecutive frames from `start`, wrapping around the end; a 1-frame tensor gives frame 0."""
    n = tensor.shape[0]
    if start + length <= n:
        indices = torch.arange(start, start + length)
    else:
        indices = torch.cat((torch.arange(start, n), torch.arange(0, (start + length) % n)))
    if n == 1:
        indices = torch.zeros(1, dtype=torch.int64)
    if return_indices:
        return indices
    return tensor[indices.to(tensor.device)]


def set_model_args(args, current_size):
    """optim.py:93-108."""
    with open(args.scaling_args, "r") as f:
        scaling = json.load(f)
    found = False
    params = {}
    for size, params in scaling.items():
        if int(size) < current_size:
            continue
        if len(str(args.gpu).split(",")) < len(str(params["gpu"]).split(",")):
            continue
        found = True
        break
    if not found:
        print("Warning: no model configuration found for this size, out of memory error is likely...")
    for key, param in params.items():
        args.__dict__[key] = param


def lbfgs_updates(num_iters: int) -> int:
    """Parameter updates of ONE torch.optim.LBFGS.step() with max_iter = num_iters as optim.py:180-191 constructs it: the
    loop also stops after max_eval = max_iter * 5 // 4 closure evaluations (torch's default, not overridden by the
    reference) -- one evaluation before the loop and one per update -- so 2 / 3 iterations give 1 / 2 updates and every
    other count gives num_iters."""
    return min(num_iters, max(num_iters * 5 // 4 - 1, 1))


class PixelOptimizer:
    """The Adam / L-BFGS pixel update of optim.py:180-196 as device kernels with device-resident state."""

    def __init__(self, pastiche: torch.Tensor, kind: str, lr: float = 1.0, history: int = 100,
                 tolerance_change: float = -1.0):
        self.lib = _lib.load()
        self.p = pastiche
        self.kind = kind
        self.lr = float(lr)
        self.step_count = 0
        self._state = None
        n = pastiche.numel()
        if kind == "adam":
            self.m = torch.zeros_like(pastiche)
            self.v = torch.zeros_like(pastiche)
            self.step_dev = torch.zeros(1, dtype=torch.int32, device=pastiche.device)  # device-resident step number
        elif kind == "lbfgs":
            self._state = C.c_void_p()
            with torch.cuda.device(pastiche.device):
                _lib.check(self.lib.maua_lbfgs_create(C.c_long(n), int(history), C.c_float(1.0), C.c_float(tolerance_change),
                                                      C.byref(self._state)), "maua_lbfgs_create")
        else:
            raise ValueError(f"unknown optimizer {kind!r}")

    def step(self, grad: torch.Tensor) -> None:
        self.step_count += 1
        with torch.cuda.device(self.p.device):
            if self.kind == "adam":
                self.step_dev.add_(1)  # on the stream: a graph replay advances it too
                _lib.check(self.lib.maua_adam_step_dev(_lib.ptr(self.p), _lib.ptr(grad), _lib.ptr(self.m), _lib.ptr(self.v),
                                                       C.c_long(self.p.numel()), C.c_float(self.lr), C.c_float(0.9),
                                                       C.c_float(0.999), C.c_float(1e-8), _lib.ptr(self.step_dev),
                                                       _lib.stream_ptr()), "maua_adam_step_dev")
            else:
                _lib.check(self.lib.maua_lbfgs_step(self._state, _lib.ptr(self.p), _lib.ptr(grad), _lib.stream_ptr()),
                           "maua_lbfgs_step")

    def reset(self) -> None:
        """Back to the freshly constructed optimizer (the reference builds a new torch.optim object per optimize() call,
        optim.py:180-196) without re-allocating the state: Adam moments / step counter zeroed, L-BFGS history emptied."""
        self.step_count = 0
        with torch.cuda.device(self.p.device):
            if self.kind == "adam":
                self.m.zero_()
                self.v.zero_()
                self.step_dev.zero_()
            else:
                _lib.check(self.lib.maua_lbfgs_reset(self._state, _lib.stream_ptr()), "maua_lbfgs_reset")

    def close(self):
        if self._state:
            self.lib.maua_lbfgs_destroy(self._state)
            self._state = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_CAPTURE_STREAMS = {}


def _capture_stream(dev) -> "torch.cuda.Stream":
    key = torch.device(dev).index
    if key not in _CAPTURE_STREAMS:
        _CAPTURE_STREAMS[key] = torch.cuda.Stream(dev)
    return _CAPTURE_STREAMS[key]


class GraphedIteration:
    """One whole reference iteration -- feval (optim.py:201-221) + pixel update (optim.py:240) -- captured once into a CUDA
    graph and replayed: ~65 kernel launches become one graph launch, which removes the host launch cost and the gaps
    between kernels (what bounds the iteration at <= 512^2).  Everything the iteration touches is device-resident and
    address-stable (plan arena, targets, optimizer state, the step number), so a replay is exactly the eager iteration.
    Falls back to eager launches when capture is impossible (layer-wise split over several devices, MAUA_NO_GRAPH=1).
    """

    def __init__(self, net, pastiche: torch.Tensor, opt: "PixelOptimizer", up: torch.Tensor, warmup: int = 2):
        import os

        self.net, self.pastiche, self.opt, self.up = net, pastiche, opt, up
        self.graph = None
        self.signature = None
        self.eager_calls = 0
        self.zero_frames = None
        self.enabled = net.n_stages == 1 and os.environ.get("MAUA_NO_GRAPH", "0") != "1"
        self.warmup = warmup  # eager iterations first: workspaces get sized, the L-BFGS "first call" path is taken eagerly

    def rearm(self, net, signature) -> None:
        """Start another optimisation with the same buffers: the warm-up iterations run eagerly again (the L-BFGS first-step
        path is host-selected), and the captured graph is kept only if it would launch exactly the same kernels."""
        self.net = net
        self.eager_calls = 0
        if signature != self.signature:
            self.graph = None
        self.signature = signature
        self.enabled = net.n_stages == 1 and __import__("os").environ.get("MAUA_NO_GRAPH", "0") != "1"

    def _eager(self):
        self.net._forward_plan(self.pastiche, keep=True)
        grad = self.net._backward_plan(self.up)
        if self.zero_frames is not None:
            # optim.py:216-219 (img_vid): frames already styled by the previous window get no gradient
            front, end = self.zero_frames
            grad[:front] = 0
            if end > 0:
                grad[-end:] = 0
        self.opt.step(grad)

    def __call__(self):
        if not self.enabled or self.eager_calls < self.warmup:
            self._eager()
            self.eager_calls += 1
            return
        if self.graph is None:
            dev = self.pastiche.device
            g = torch.cuda.CUDAGraph()
            count = self.opt.step_count
            try:
                # capture_begin / capture_end on a side stream instead of the torch.cuda.graph() context manager: that one
                # also runs gc.collect() and torch.cuda.empty_cache() on entry, which costs 20-260 ms per capture (measured:
                # tools/dbg_small_scale.py) -- more than 200 iterations at 256^2 -- every time a scale starts
                side = _capture_stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    g.capture_begin()
                    try:
                        self._eager()
                    finally:
                        g.capture_end()
                torch.cuda.current_stream(dev).wait_stream(side)
            except Exception as e:  # noqa: BLE001
                # capture unsupported in this environment: stay eager (the capture launched nothing) -- and say so,
                # because eager launches are several times slower at <= 512^2
                import warnings

                warnings.warn(f"maua_style_b200: CUDA-graph capture of the iteration failed ({type(e).__name__}: {e}); "
                              "falling back to eager kernel launches", RuntimeWarning)
                self.enabled = False
                self.opt.step_count = count
                torch.cuda.synchronize(dev)
                self._eager()
                return
            self.opt.step_count = count  # the capture itself does not execute
            self.graph = g
        self.graph.replay()
        self.opt.step_count += 1


class HostPipelinedIteration:
    """An iteration driven from HOST buffers without stalling the GPU: every `submit(host_image)` copies that step's input
    image host->device from pinned memory, runs one fused iteration on it and copies the updated image + total loss back
    to pinned host memory -- with the copies on their own streams, double-buffered, so that step i's device->host read
    and step i+1's host->device copy run under step i+1's / step i's kernels.  `submit` returns the result of the
    PREVIOUS step (already on the host); `drain()` returns the last one.  Used by bench.py's end-to-end number and by
    callers that stream frames through one network."""

    def __init__(self, iteration: "GraphedIteration"):
        self.it = iteration
        p = iteration.pastiche
        dev = p.device
        self.dev = dev
        self.copy_in, self.copy_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.dev_in = [torch.empty_like(p) for _ in range(2)]
        self.dev_out = [torch.empty_like(p) for _ in range(2)]
        self.dev_loss = [torch.zeros(1, device=dev) for _ in range(2)]
        self.host_out = [torch.empty_like(p, device="cpu").pin_memory() for _ in range(2)]
        self.host_loss = [torch.empty(1).pin_memory() for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_consumed = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.n = 0
        self.pending = None
        self.bytes_in = p.numel() * 4
        self.bytes_out = p.numel() * 4 + 4

    def submit(self, host_image: torch.Tensor):
        b = self.n % 2
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_in):
            if self.n >= 2:
                self.copy_in.wait_event(self.ev_consumed[b])      # the step two back has read this staging buffer
            self.dev_in[b].copy_(host_image, non_blocking=True)    # H2D of THIS step's input
            self.ev_in[b].record(self.copy_in)
        main.wait_event(self.ev_in[b])
        self.it.pastiche.copy_(self.dev_in[b])
        self.ev_consumed[b].record(main)
        self.it()
        self.dev_out[b].copy_(self.it.pastiche)
        torch.sum(self.it.net._loss_vec, dim=0, keepdim=True, out=self.dev_loss[b])
        self.ev_done[b].record(main)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(self.ev_done[b])
            self.host_out[b].copy_(self.dev_out[b], non_blocking=True)   # D2H of this step's result
            self.host_loss[b].copy_(self.dev_loss[b], non_blocking=True)
            self.ev_out[b].record(self.copy_out)
        prev, self.pending = self.pending, b
        self.n += 1
        if prev is None:
            return None
        self.ev_out[prev].synchronize()                            # the previous step's result is on the host now
        return self.host_out[prev], self.host_loss[prev]

    def drain(self):
        if self.pending is None:
            return None
        b, self.pending = self.pending, None
        self.ev_out[b].synchronize()
        return self.host_out[b], self.host_loss[b]


def feval(net, pastiche: torch.Tensor, ones: Optional[torch.Tensor] = None):
    """optim.py:201-221 without the host syncs: returns (loss vector on the device, pastiche gradient)."""
    net._forward_plan(pastiche, keep=True)
    live = net._live_slots()
    up = torch.zeros(net._n_slots, device=net.device)
    up[live] = 1.0
    grad = net._backward_plan(up)
    return net._loss_vec, grad


def optimize(content, styles, init, num_iters, args, net=None, losses=None):
    """optim.py:111-255 for transfer types img_img / vid_img (one window, batch 1): returns the CPU tensor the reference
    returns (optim.py:249)."""
    return optimize_device(content, styles, init, num_iters, args, net, losses).cpu()


def optimize_device(content, styles, init, num_iters, args, net=None, losses=None):
    """`optimize` without the final device->host copy: the result stays in HBM for the next scale / frame
    (maua_style_b200/style.py keeps the whole multi-resolution schedule on the device)."""
    if "_vid" in getattr(args, "transfer_type", "img_img"):
        return _optimize_windows(content, styles, init, num_iters, args, net, losses)
    tr = _Trace(f"optimize {tuple(init.shape[2:])}")
    if net is None or losses is None:
        set_model_args(args, max(*init.shape))
        net, losses = models.load_model(args)
    device = net.device
    tr.mark(f"load_model (plan-core cache: {models.cache_stats['hits']} hits / {models.cache_stats['misses']} misses so far)")

    import os

    if os.environ.get("MAUA_NO_LOOP_CACHE", "0") != "1":
        # content / temporal targets of the same shape are re-captured into their existing buffers, so that the iteration
        # captured for the previous image of this size can be replayed for this one (see _LoopState)
        net.reuse_target_buffers = True
    set_content_targets(net, content.to(device, torch.float32), args)
    set_style_targets(net, styles, args)  # (the network moves host tensors itself; identity is kept for the target cache)
    for mod in losses:
        mod.mode = "loss"
    tr.mark("target capture")

    # optim.py:176-178 (only once, strengths are not reset)
    if getattr(args, "normalize_weights", False):
        for i in net.content_losses + net.style_losses + net.temporal_losses:
            i.strength = i.strength / max(i.target.size())

    if args.optimizer == "lbfgs":
        # optim.py:180-186 passes tolerance_grad = -1 (never stops on the gradient norm); the device L-BFGS has no such
        # test, so a value that could fire is refused instead of silently ignored
        if float(getattr(args, "lbfgs_tolerance_grad", -1)) >= 0:
            raise NotImplementedError("maua_style_b200: lbfgs_tolerance_grad >= 0 is not supported (the reference always "
                                      "passes -1, optim.py:184); use the default")
        evals = lbfgs_updates(num_iters)  # one step() = num_iters closure evaluations and updates
    elif args.optimizer == "adam":
        evals = num_iters + 1  # optim.py:240 `while i[0] <= iters`
    else:
        raise ValueError(f"unknown optimizer {args.optimizer!r}")

    live = net._live_slots()  # modules in loss mode with a target (a shape mismatch just contributes 0, loss.py:44)
    state = _loop_state(net, init, args, live)
    pastiche, opt, iteration = state.pastiche, state.opt, state.iteration
    tr.mark("loop state")

    print_iter = int(getattr(args, "print_iter", 0) or 0)
    save_iter = int(getattr(args, "save_iter", 0) or 0)
    for it in range(1, evals + 1):
        want_print = print_iter > 0 and it % print_iter == 0 and getattr(args, "verbose", False)
        want_save = save_iter > 0 and (it % save_iter == 0 or it == num_iters)
        if want_save:
            # optim.py:230-236 saves the image the closure was evaluated on, i.e. before this iteration's update
            _save_intermediate(pastiche, args, it, num_iters)
        iteration()
        if tr.fine and it % 100 == 0:
            tr.mark(f"iterations {it - 99}..{it}")
        if want_print:
            total = float(net._loss_vec[live].sum())  # the losses of the image before the update, like optim.py:228-229
            print(f"Iteration {it} / {args.num_iters}, Loss: {total}")
    tr.mark(f"{evals} iterations")
    for mod in losses:
        mod.loss = 0
    if state.cached:
        iteration.net = None     # the cached state must not keep the network (and through it the plan core's owner) alive
        return pastiche.clone()  # the buffer belongs to the cached loop state and is overwritten by the next call
    opt.close()
    return pastiche


def _optimize_windows(content, styles, init, num_iters, args, net=None, losses=None):
    """optim.py:113-125, :149-170, :216-219, :242-247 -- transfer type img_vid: the pastiche is a video [T, 3, H, W] optimised
    in overlapping windows of `gram_frame_window` frames (a batch of B frames through the network, window.py); window
    starts are spread evenly over the pastiche and over every style video; frames a previous window already styled get
    no gradient; every window starts a fresh optimizer."""
    gfw = int(args.gram_frame_window)
    afw = int(getattr(args, "avg_frame_window", -1))
    clips = [init] + list(styles)
    num_windows = math.ceil(init.shape[0] / gfw)
    framestep = [(c.shape[0] - gfw / 2) / num_windows for c in clips]
    windows = [[math.ceil(framestep[idx] * n) for n in range(num_windows + 1)] if clips[idx].shape[0] != 1
               else [0] * (num_windows + 1) for idx in range(len(clips))]
    if net is None or losses is None:
        set_model_args(args, max(*init.shape))
        net, losses = models.load_model(args)
    device = net.device
    if args.optimizer == "lbfgs":
        if float(getattr(args, "lbfgs_tolerance_grad", -1)) >= 0:
            raise NotImplementedError("maua_style_b200: lbfgs_tolerance_grad >= 0 is not supported")
        evals = lbfgs_updates(num_iters)
    elif args.optimizer == "adam":
        evals = num_iters + 1
    else:
        raise ValueError(f"unknown optimizer {args.optimizer!r}")

    set_content_targets(net, content.to(device, torch.float32), args)
    styles_dev = [s.to(device, torch.float32) for s in styles]
    if afw == -1:
        set_style_video_targets(net, styles_dev, args)
        for mod in losses:
            mod.mode = "loss"
    output = init.detach().to(device, torch.float32).clone()
    T = output.shape[0]
    for w, window_start in enumerate(windows[0]):
        front_overlap = windows[0][w - 1] + gfw - window_start
        end_overlap = (window_start + gfw) % T if window_start + gfw >= T else 0
        indices = wrapping_slice(output, window_start, gfw, return_indices=True).to(device)
        if afw != -1:
            current_styles = [wrapping_slice(style, windows[num + 1][w], afw) for num, style in enumerate(styles_dev)]
            set_style_video_targets(net, current_styles, args)
            for mod in losses:
                mod.mode = "loss"
        pastiche = output[indices].contiguous()
        if w == 0 and getattr(args, "normalize_weights", False):  # optim.py:176-178
            for i in net.content_losses + net.style_losses + net.temporal_losses:
                i.strength = i.strength / max(i.target.size())
        if args.optimizer == "lbfgs":
            opt = PixelOptimizer(pastiche, "lbfgs", history=int(getattr(args, "lbfgs_num_correction", 100)),
                                 tolerance_change=float(getattr(args, "lbfgs_tolerance_change", -1)))
        else:
            opt = PixelOptimizer(pastiche, "adam", lr=float(getattr(args, "learning_rate", 1.0)))
        live = net._live_slots()
        up = torch.zeros(net._n_slots, device=device)
        up[live] = 1.0
        iteration = GraphedIteration(net, pastiche, opt, up)
        if w != 0:
            iteration.zero_frames = (front_overlap, end_overlap)
        for _ in range(evals):
            iteration()
        output[indices] = pastiche  # optim.py:244
        opt.close()
    for mod in losses:
        mod.loss = 0
    return output


class _LoopState:
    """What one optimisation loop needs besides the network: the device-resident pastiche, the optimizer state and the
    captured iteration.  Kept on the plan core per (image size, optimizer settings) so that the next image / frame /
    call of the same size re-uses the buffers (an L-BFGS history at 1024^2 is 2.5 GB of cudaMalloc + cudaFree per call
    otherwise) and -- when the kernel arguments are identical -- the captured CUDA graph."""

    def __init__(self, key, pastiche, opt, iteration, cached):
        self.key, self.pastiche, self.opt, self.iteration, self.cached = key, pastiche, opt, iteration, cached


_LOOP_STATES_MAX = 4


def _loop_state(net, init, args, live) -> _LoopState:
    import os

    device = net.device
    kind = args.optimizer
    hist = int(getattr(args, "lbfgs_num_correction", 100))
    tol = float(getattr(args, "lbfgs_tolerance_change", -1))
    lr = float(getattr(args, "learning_rate", 1.0)) if kind == "adam" else 1.0
    key = (tuple(init.shape), kind, hist, tol, lr)
    cache = getattr(getattr(net, "_core", None), "loop_states", None)
    use_cache = cache is not None and os.environ.get("MAUA_NO_LOOP_CACHE", "0") != "1"
    H, W = int(init.shape[2]), int(init.shape[3])
    signature = (net.io_signature(H, W), tuple(live))
    state = cache.get(key) if use_cache else None
    if state is not None:
        cache.move_to_end(key)
        state.pastiche.copy_(init.detach().to(device, torch.float32))  # optim.py:173
        state.opt.reset()
        state.iteration.up.zero_()
        state.iteration.up[live] = 1.0
        state.iteration.rearm(net, signature)
        return state
    # optim.py:173: the pastiche lives on the device for the whole optimisation
    pastiche = init.detach().to(device, torch.float32).contiguous().clone()
    if kind == "lbfgs":
        opt = PixelOptimizer(pastiche, "lbfgs", history=hist, tolerance_change=tol)
    else:
        opt = PixelOptimizer(pastiche, "adam", lr=lr)
    up = torch.zeros(net._n_slots, device=device)
    up[live] = 1.0
    iteration = GraphedIteration(net, pastiche, opt, up)
    iteration.signature = signature
    state = _LoopState(key, pastiche, opt, iteration, use_cache)
    if use_cache:
        cache[key] = state
        while len(cache) > _LOOP_STATES_MAX:
            _, old = cache.popitem(last=False)
            old.opt.close()
    return state


def _save_intermediate(pastiche, args, it, num_iters):
    """optim.py:230-236 hands the image to load.save_tensor_to_file (host I/O, outside the hot path)."""
    saver = getattr(args, "save_fn", None)
    if saver is None:
        try:
            import load  # the reference's I/O module, when running inside the reference tree

            saver = lambda img, a, i, size: load.save_tensor_to_file(img, a, i, size)
        except Exception:
            return
    saver(pastiche.detach().cpu(), args, it if it != num_iters else None, pastiche.size(3))
