"""The multi-resolution drivers of the reference's style.py, with everything between the decoded input images and the
encoded output image resident on the B200.

The reference's own `style.py` runs unchanged on top of `maua_style_b200.{loss,models,optim}`; this
module is what it looks like when the steps *around* `optim.optimize` stop bouncing through the host as well:

  reference (style.py:22-73, per scale)                       here
  ------------------------------------------------------------------------------------------------------------------
  F.interpolate(content / styles / pastiche) on the CPU       image_ops.interpolate on the device (bit-identical)
  optim.optimize: set_model_args + load_model every scale     one plan core re-used across scales (models.build_net cache)
      (torch.load + deepcopy + .cuda(), optim.py:128-129)
  pastiche.cpu() -> match_histogram -> next scale             pastiche stays in HBM (optim.optimize_device)
  load.save_tensor_to_file: fp32 D2H, deprocess on the CPU    deprocess on the device, 3 B/pixel D2H (image_ops.deprocess)

`img_img_tensors` / `vid_img_tensors` / `img_vid_tensors` / `stylize_frame` work on tensors (what benchmarks and the sharded runner call); `img_img(args)` is
the file-level entry with the reference's argument names.  Histogram matching (utils.match_histogram, style.py:24, :67,
:71) runs on the device when `args.match_histograms` is set (image_ops.match_histogram: the style images' colour
moments are taken once, each call is two passes over the pastiche).  Note that on torch >= 2 the reference's own call
silently does nothing because `th.symeig` no longer exists; with the flag on, this driver
does what the reference did under its pinned torch 1.8.1.
"""
from __future__ import annotations

import math
import os
from typing import Callable, List, Optional, Sequence

import torch

from . import _lib, image_ops, models, optim


def _device(args) -> torch.device:
    return models._device_from_args(args)


def scale_schedule(content_hw, style_hws, image_sizes: Sequence[int], style_scale: float = 1.0):
    """style.py:36-50: per scale the content scale factor / size and every style image's scale factor / size."""
    out = []
    for size in image_sizes:
        cs = size / max(*content_hw)
        ch, cw = image_ops.interp_out_size(content_hw[0], cs), image_ops.interp_out_size(content_hw[1], cs)
        styles = []
        for sh, sw in style_hws:
            ss = math.sqrt((ch * cw) / (sw * sh)) * style_scale
            styles.append((ss, (image_ops.interp_out_size(sh, ss), image_ops.interp_out_size(sw, ss))))
        out.append({"size": size, "content_scale": cs, "content_hw": (ch, cw), "styles": styles})
    return out


def img_img_tensors(content_big: torch.Tensor, styles_big: Sequence[torch.Tensor], args, init_image: Optional[torch.Tensor] = None,
                    on_scale: Optional[Callable[[int, torch.Tensor], None]] = None) -> List[torch.Tensor]:
    """style.py:22-73 on tensors.  `content_big` / `styles_big` / `init_image`: [1,3,H,W] preprocessed images (BGR, 0-255,
    mean-subtracted; host or device).  Uses args.image_sizes, args.num_iters, args.init ("content" | "random" | anything
    else = `init_image`), args.style_scale and everything `optim.optimize` reads.  Returns the pastiche of every scale as
    device tensors; `on_scale(size, pastiche)` is called after each scale (e.g. to save it)."""
    dev = _device(args)
    with torch.cuda.device(dev):
        content_big = content_big.to(dev, torch.float32).contiguous()
        styles_big = [s.to(dev, torch.float32).contiguous() for s in styles_big]
        # style.py:24 -- the content image takes the colour statistics of the style images (moments taken once)
        hist = getattr(args, "match_histograms", False)
        style_moments = None
        if hist:
            style_moments = torch.stack([image_ops.image_moments(s) for s in styles_big])
            content_big = image_ops.match_histogram(content_big, None, mode=hist, source_moments=style_moments)
        init = getattr(args, "init", "content")
        pastiche = None
        if init not in ("content", "random"):
            if init_image is None:
                raise ValueError("args.init names an image: pass it as init_image")
            pastiche = init_image.to(dev, torch.float32).contiguous()
        outs = []
        for current_size, num_iters in zip(args.image_sizes, args.num_iters):
            # scale content image (style.py:36-41)
            content_scale = current_size / max(*content_big.shape[-2:])
            content_image = image_ops.interpolate(content_big, scale_factor=content_scale)
            # scale style images (style.py:43-50)
            content_area = content_image.shape[2] * content_image.shape[3]
            style_images = []
            for img in styles_big:
                style_scale = math.sqrt(content_area / (img.size(3) * img.size(2))) * getattr(args, "style_scale", 1.0)
                style_images.append(image_ops.interpolate(img, scale_factor=style_scale))
            # initialise the pastiche (style.py:52-66)
            if init == "random" and pastiche is None:
                H, W = content_image.shape[2:]
                pastiche = torch.randn(1, 3, H, W).mul(0.001).to(dev)
            elif init == "content" and pastiche is None:
                pastiche = image_ops.interpolate(content_big, size=tuple(content_image.shape[2:]))
            else:
                pastiche = image_ops.interpolate(pastiche, size=tuple(content_image.shape[2:]))
            if hist:  # style.py:67
                pastiche = image_ops.match_histogram(pastiche, None, mode=hist, source_moments=style_moments, out=pastiche)
            # style.py:69 -- per-size model args + (cached) model + targets + the optimisation loop, result stays in HBM
            pastiche = optim.optimize_device(content_image, style_images, pastiche, num_iters, args)
            if hist:  # style.py:71
                pastiche = image_ops.match_histogram(pastiche, None, mode=hist, source_moments=style_moments, out=pastiche)
            outs.append(pastiche)
            if on_scale is not None:
                on_scale(current_size, pastiche)
        return outs


def stylize_frame(net, losses, content_frame: torch.Tensor, style_images: Sequence[torch.Tensor], args, num_iters: int,
                  prev_pastiche: Optional[torch.Tensor] = None, flow_grid: Optional[torch.Tensor] = None,
                  reliable_flow: Optional[torch.Tensor] = None, blend_image: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One frame of vid_img (style.py:273-297) on the device: warp the previous frame's pastiche along the flow
    (`F.grid_sample(..., padding_mode="border")`, :276), resize the flow-reliability map (:278-282), capture the temporal
    target (:284), blend the initialisation (:286) and optimise (:288-290).  `net, losses` come from `models.load_model`
    once per scale (:176-177); the style targets are captured for the first frame only (optim.set_style_targets cache).
    Without `prev_pastiche` the frame starts from the content frame (style.py:229-230)."""
    dev = net.device
    with torch.cuda.device(dev):
        content_frame = content_frame.to(dev, torch.float32).contiguous()
        if prev_pastiche is None:
            pastiche = content_frame.clone()
        else:
            pastiche = prev_pastiche.to(dev, torch.float32).contiguous()
            if tuple(pastiche.shape[2:]) != tuple(content_frame.shape[2:]):
                pastiche = image_ops.interpolate(pastiche, size=tuple(content_frame.shape[2:]))  # style.py:242-244
            if flow_grid is not None:
                warp_image = image_ops.grid_sample(pastiche, flow_grid.to(dev))
                if reliable_flow is not None:
                    reliable_flow = image_ops.interpolate(reliable_flow.to(dev, torch.float32), size=tuple(pastiche.shape[2:]))
                optim.set_temporal_targets(net, warp_image, warp_weights=reliable_flow, args=args)
            if blend_image is not None:
                tb = float(getattr(args, "temporal_blend", 0.5))
                blend_image = blend_image.to(dev, torch.float32).contiguous()
                if tuple(blend_image.shape[2:]) != tuple(pastiche.shape[2:]):
                    blend_image = image_ops.interpolate(blend_image, size=tuple(pastiche.shape[2:]))  # style.py:252-254
                pastiche = image_ops.blend(blend_image, pastiche, 1.0 - tb, tb)
        return optim.optimize_device(content_frame, style_images, pastiche, num_iters, args, net, losses)


def vid_img_pairs(order: Sequence[int], loop: bool = False):
    """style.py:192-194: the (previous frame, this frame) pairs of one pass over the frame list `order`.  Without --loop every
    frame is `this frame` once and the first frame comes last (it follows the last one); with --loop the first frames are
    styled a second time so that the end of the clip meets its start."""
    order = list(order)
    return list(zip(order + order[: 11 if loop else 1], order[1:] + order[: 10 if loop else 1]))


def vid_img_tensors(frames: Sequence[torch.Tensor], styles_big: Sequence[torch.Tensor], args,
                    flows: Callable[[str, int, int], tuple],
                    on_frame: Optional[Callable[[int, int, int, torch.Tensor], None]] = None,
                    owned: Optional[Sequence[int]] = None,
                    exchange: Optional[Callable[[dict], dict]] = None) -> dict:
    """style.py:145-300 on tensors: every scale x `args.passes_per_scale` passes (forward, then backward over the reversed
    frame list, :300) x every frame, with everything between the decoded frames and the encoded results on the device.

    frames ...... the decoded clip, [1,3,H,W] preprocessed images (load.preprocess layout; host or device)
    styles_big .. the style images, same layout
    flows ....... `flows(direction, prev_index, this_index) -> (flow, reliable)`: what the reference reads from
                  `<work_dir>/flow/<direction>_<prev>_<this>.flo` / `.png` (:226-227, :273-274, :278-279) -- the flow field [h,w,2] after
                  the host-side normalisation + blur of load.py:201-206 (`read_flo`), and the flow-reliability map [1,1,h,w]
                  in [0,1] (load.py:217-218).  Estimating the flow is not part of this path: the fields are inputs.
    on_frame .... called as on_frame(size, pass (1-based), frame index, uint8 [H,W,3] RGB device tensor) for every result,
                  e.g. to encode `<size>/<pass>_<frame>.png` (:197, :294-297)

    The reference hands results from one pass / scale to the next through those PNG files (:232-271); here they are the
    returned dict {(size, pass, frame index): uint8 image in HBM} and never leave the device, but they keep the 8-bit
    quantisation the files impose, so the frames equal the reference's.  Within a pass the previous frame's fp32 result is
    carried over directly (:292).  Reads args.image_sizes, num_iters, passes_per_scale, init ("random" | "prev_warp" | else
    the content frame), temporal_blend, loop, style_scale, match_histograms and everything `optim.optimize` reads; calls
    `optim.set_model_args(args, size)` per scale like the reference does (:176), i.e. it updates `args` in place.

    Sharding over GPUs (shard.stylize_video): `owned` = the frame indices this process styles (default: all), `exchange` =
    called after every pass with the frames this process produced in it, returns the frames of all processes.  A frame that
    another process styles breaks the chain of carried-over results exactly like a frame whose PNG already exists does in the
    reference's resume path (:198-200): the next owned frame starts from the stored result of its predecessor (:232-271)."""
    import random

    dev = _device(args)
    n = len(frames)
    if n < 2:
        raise ValueError("vid_img needs at least two frames")
    passes = int(getattr(args, "passes_per_scale", 1))
    tb = float(getattr(args, "temporal_blend", 0.5))
    loop = bool(getattr(args, "loop", False))
    hist = getattr(args, "match_histograms", False)
    init = getattr(args, "init", "prev_warp")
    store = {}
    mine = None if owned is None else set(int(i) for i in owned)
    if mine is not None and (loop or init == "random"):
        raise NotImplementedError("maua_style_b200: sharded vid_img draws nothing at random (no --loop rotation, no random init): "
                                  "every process must walk the same schedule")
    with torch.cuda.device(dev):
        frames = [f.to(dev, torch.float32).contiguous() for f in frames]
        styles_big = [s.to(dev, torch.float32).contiguous() for s in styles_big]
        H, W = (int(v) for v in frames[0].shape[-2:])
        moments = torch.stack([image_ops.image_moments(styles_big[0])]) if hist else None  # :212-214, :292: the first style image

        def matched(img):
            return image_ops.match_histogram(img, None, mode=hist, source_moments=moments) if hist else img

        def stored(key):  # load.preprocess(<png>) of a frame an earlier pass / scale wrote
            return image_ops.preprocess(store[key], dev)

        order = list(range(n))
        prev_size = None
        for size_n, (current_size, num_iters) in enumerate(zip(args.image_sizes, args.num_iters)):
            content_scale = current_size / max(H, W)
            # scale style images (:167-174; the area comes from the un-rounded scale factor, unlike img_img)
            content_area = content_scale ** 2 * H * W
            style_images = []
            for img in styles_big:
                style_scale = math.sqrt(content_area / (img.size(3) * img.size(2))) * getattr(args, "style_scale", 1.0)
                style_images.append(image_ops.interpolate(img, scale_factor=style_scale))
            optim.set_model_args(args, current_size)  # :176-177: one network per scale, every frame re-uses it
            net, losses = models.load_model(args)
            for pass_n in range(passes):
                pastiche = None
                if loop:  # :183-185
                    start = random.randrange(0, n - 1)
                    order = order[start:] + order[:start]
                direction = "forward" if pass_n % 2 == 0 else "backward"  # :215
                fresh = {}
                for k, (prev_f, this_f) in enumerate(vid_img_pairs(order, loop)):
                    if mine is not None and this_f not in mine:
                        pastiche = None  # someone else's frame: the chain of carried-over results ends here
                        continue
                    content_frame = matched(image_ops.interpolate(frames[this_f], scale_factor=content_scale))
                    if size_n == 0 and pass_n == 0:  # :220-230
                        if init == "random":
                            pastiche = torch.randn(content_frame.size()).mul(0.001).to(dev)
                        elif init == "prev_warp":
                            if pastiche is None:
                                pastiche = matched(image_ops.interpolate(frames[prev_f], scale_factor=content_scale))
                            flow, _ = flows(direction, prev_f, this_f)
                            pastiche = image_ops.grid_sample(pastiche, image_ops.flow_warp_grid(flow, tuple(pastiche.shape[2:])))
                        else:
                            pastiche = content_frame.clone()
                    else:  # :231-286
                        if pass_n == 0:  # the last pass of the previous scale (:232-254) ...
                            src = (prev_size, passes) if k <= n else (current_size, pass_n + 1)
                        else:            # ... or the previous pass of this scale (:255-271)
                            src = (current_size, pass_n) if k <= n else (current_size, pass_n + 1)
                        hw = tuple(content_frame.shape[2:])
                        if pastiche is None:
                            pastiche = stored(src + (prev_f,))
                            if tuple(pastiche.shape[2:]) != hw:
                                pastiche = image_ops.interpolate(pastiche, size=hw)
                        blend_image = stored(src + (this_f,))
                        if tuple(blend_image.shape[2:]) != hw:
                            blend_image = image_ops.interpolate(blend_image, size=hw)
                        flow, reliable = flows(direction, prev_f, this_f)
                        warp_image = image_ops.grid_sample(pastiche, image_ops.flow_warp_grid(flow, hw))  # :273-276
                        reliable = image_ops.interpolate(reliable.to(dev, torch.float32), size=hw)        # :278-282
                        optim.set_temporal_targets(net, warp_image, warp_weights=reliable, args=args)     # :284
                        pastiche = image_ops.blend(blend_image, pastiche, 1.0 - tb, tb)                    # :286
                    out = optim.optimize_device(content_frame, style_images, pastiche, num_iters // passes, args, net, losses)
                    pastiche = matched(out)  # :292
                    u8 = image_ops.deprocess_u8(pastiche)
                    store[(current_size, pass_n + 1, this_f)] = fresh[(current_size, pass_n + 1, this_f)] = u8
                    if on_frame is not None:
                        on_frame(current_size, pass_n + 1, this_f, u8)
                if exchange is not None:
                    store.update(exchange(fresh))  # the next pass / scale reads its neighbours' frames (:232-271)
                order = list(reversed(order))  # :300
            prev_size = current_size
    return store


def temporal_blur(video: torch.Tensor, sigma: float) -> torch.Tensor:
    """`ndi.gaussian_filter(video, [sigma, 0, 0, 0], mode="wrap")` (style.py:137-138) without leaving the device: a normalised
    Gaussian of radius int(4 sigma + 0.5) along the frame axis with periodic extension, accumulated in float64 like scipy."""
    r = int(4.0 * float(sigma) + 0.5)
    ks = list(range(-r, r + 1))
    w = [math.exp(-0.5 / (sigma * sigma) * k * k) for k in ks]
    total = sum(w)
    v = video.to(torch.float64)
    out = torch.zeros_like(v)
    for k, wk in zip(ks, w):
        out += (wk / total) * torch.roll(v, -k, dims=0)  # out[t] += w[k] * v[(t + k) mod T]
    return out.to(torch.float32)


def initial_video(content_big: torch.Tensor, video_length: int, init: str = "content") -> torch.Tensor:
    """style.py:93-103: the initial pastiche video of img_vid for init "random" / "content" -- seeded noise (torch's global RNG,
    like the reference) blurred over time and space with scipy on the host, once per job.  Returns a host tensor [T,3,H,W]."""
    import scipy.ndimage as ndi

    H, W = (int(v) for v in content_big.shape[-2:])
    if init == "random":
        pastiche = torch.randn((video_length, 3, H, W)) * 255
        return torch.from_numpy(ndi.gaussian_filter(pastiche.numpy(), [video_length, 0, H / 32, W / 32], mode="wrap"))
    if init != "content":
        raise ValueError("init names a video: pass it as init_video")
    pastiche = content_big.detach().to("cpu", torch.float32).clone().repeat([video_length, 1, 1, 1])
    pastiche += torch.randn((video_length, 3, H, W)) * 255
    return torch.from_numpy(ndi.gaussian_filter(pastiche.numpy(), [video_length, 0, 4, 4], mode="wrap"))


def img_vid_tensors(content_big: torch.Tensor, style_videos_big: Sequence[torch.Tensor], args,
                    init_video: Optional[torch.Tensor] = None,
                    on_scale: Optional[Callable[[int, torch.Tensor], None]] = None) -> List[torch.Tensor]:
    """style.py:76-142 on tensors: the pastiche is a video [T,3,H,W] that `optim.optimize` styles in overlapping frame windows
    (one window = one batch of B frames through the network, window.py); between scales the pastiche and the style clips are
    rolled by 7 frames (:134-135) so that the window seams fall elsewhere, and the pastiche is blurred over time (:137-138).
    The video stays in HBM across scales.

    content_big: [1,3,H,W]; style_videos_big: clips [T_i,3,h,w] (load.preprocess layout); `init_video`: the initial pastiche
    (default: `initial_video(content_big, T, args.init)`, :93-103; a shorter clip is repeated like :102-103).  Reads
    args.image_sizes, num_iters, gram_frame_window (one window length per scale: "18,9,7", a list, or one int), num_frames
    (-1: the longest style clip), temporal_blend, style_scale and everything `optim.optimize` reads; like the reference it
    sets `args.gram_frame_window` to the current scale's value.  Histogram matching against style *videos* (:82, :104, :139, :142)
    is refused: on torch >= 2 the reference's own call is a no-op and the device version
    handles single images."""
    if getattr(args, "match_histograms", False):
        raise NotImplementedError("maua_style_b200: match_histograms with style videos is not supported (pass --no_hist_match)")
    dev = _device(args)
    nf = int(getattr(args, "num_frames", -1))
    video_length = max(int(v.shape[0]) for v in style_videos_big) if nf == -1 else nf
    gfw = getattr(args, "gram_frame_window")
    delta_ts = [int(x) for x in gfw.split(",")] if isinstance(gfw, str) else ([int(x) for x in gfw] if isinstance(gfw, (list, tuple)) else [int(gfw)] * len(args.image_sizes))
    if len(delta_ts) < len(args.image_sizes):
        raise ValueError("gram_frame_window needs one window length per image size")
    tb = float(getattr(args, "temporal_blend", 0.0))
    if init_video is None:
        init_video = initial_video(content_big, video_length, getattr(args, "init", "content"))
    elif init_video.shape[0] != video_length:
        init_video = init_video.repeat([video_length, 1, 1, 1])  # :102-103
    with torch.cuda.device(dev):
        content_big = content_big.to(dev, torch.float32).contiguous()
        clips = [v.to(dev, torch.float32).contiguous() for v in style_videos_big]
        pastiche = init_video.to(dev, torch.float32).contiguous()
        H, W = (int(v) for v in content_big.shape[-2:])
        outs = []
        for i, (current_size, num_iters) in enumerate(zip(args.image_sizes, args.num_iters)):
            args.gram_frame_window = delta_ts[i]  # :111
            content_image = image_ops.interpolate(content_big, scale_factor=current_size / max(H, W))  # :114-116
            content_area = content_image.shape[2] * content_image.shape[3]
            style_videos = []
            for vid in clips:  # :119-125
                style_scale = math.sqrt(content_area / (vid.size(3) * vid.size(2))) * getattr(args, "style_scale", 1.0)
                style_videos.append(image_ops.interpolate(vid, scale_factor=style_scale))
            pastiche = image_ops.interpolate(pastiche, size=tuple(content_image.shape[2:]))  # :128-130
            pastiche = optim.optimize_device(content_image, style_videos, pastiche, num_iters, args)  # :132
            pastiche = torch.cat((pastiche[7:], pastiche[:7]))  # :134-135
            clips = [torch.cat((c[7:], c[:7])) for c in clips]
            if tb > 0:
                pastiche = temporal_blur(pastiche, tb)
            outs.append(pastiche)
            if on_scale is not None:
                on_scale(current_size, pastiche)
        return outs


def read_flo(path: str) -> torch.Tensor:
    """The host-side half of load.flow_warp_map (load.py:191-206): a Middlebury .flo file -> the field normalised by its own
    extent and blurred (sigma 5), [h,w,2] float32 -- the `flow` input of `vid_img_tensors` / `image_ops.flow_warp_grid`."""
    import numpy as np
    import scipy.ndimage

    with open(path, "rb") as f:
        if np.fromfile(f, np.float32, count=1)[0] != np.float32(202021.25):
            raise ValueError(f"{path}: not a .flo file (bad magic number)")
        w = int(np.fromfile(f, np.int32, count=1)[0])
        h = int(np.fromfile(f, np.int32, count=1)[0])
        flow = np.fromfile(f, np.float32, count=2 * w * h).reshape(h, w, 2).copy()
    flow[:, :, 0] /= w
    flow[:, :, 1] /= h
    return torch.from_numpy(scipy.ndimage.gaussian_filter(flow, [5, 5, 0]))


# ---------------------------------------------------------------------------------------------------------------------
# file-level entry with the reference's argument names (host I/O is PIL; decode / encode are outside the hot path)
# ---------------------------------------------------------------------------------------------------------------------
def load_image(path: str, device) -> torch.Tensor:
    """load.preprocess(path) (load.py:21-32) with the arithmetic on the device."""
    from PIL import Image

    Image.MAX_IMAGE_PIXELS = 1000000000  # load.py:15
    return image_ops.preprocess(Image.open(path).convert("RGB"), device)


def save_image(t: torch.Tensor, filename: str) -> None:
    """load.save_tensor_to_file for one image (load.py:55-75) without original_colors."""
    image_ops.deprocess(t).save(filename)


def img_img(args) -> List[torch.Tensor]:
    """style.py:22-73: args.content, args.style (list of paths), args.output (prefix), args.image_sizes, args.num_iters."""
    _lib.require_gpu()
    dev = _device(args)
    styles_big = [load_image(p, dev) for p in args.style]
    content_big = load_image(args.content, dev)
    init_image = None
    if args.init not in ("content", "random"):
        init_image = load_image(args.init, dev)
    done = {}

    def on_scale(size, pastiche):
        print("\nCurrent size {}px".format(size))
        save_image(pastiche, f"{args.output}_{size}.png")
        done[size] = True

    # resume (style.py:31-33): scales whose PNG exists are skipped and that PNG seeds the next scale
    sizes, iters = list(args.image_sizes), list(args.num_iters)
    while sizes and os.path.exists(f"{args.output}_{sizes[0]}.png"):
        init_image = load_image(f"{args.output}_{sizes[0]}.png", dev)
        sizes, iters = sizes[1:], iters[1:]
    if not sizes:
        return []
    import copy

    a = copy.copy(args)
    a.image_sizes, a.num_iters = sizes, iters
    if init_image is not None:
        a.init = "image"
    return img_img_tensors(content_big, styles_big, a, init_image=init_image, on_scale=on_scale)


def _name(path: str) -> str:
    """load.name (load.py:99-100)."""
    return str(path).split("/")[-1].split(".")[0]


def vid_img(args, frames: Optional[Sequence[str]] = None) -> dict:
    """style.py:145-300 at file level, on the files the reference's own preparation stage leaves in
    `<output_dir>/<content>_<styles>/` (load.process_content_video, load.py:141-188: ffmpeg frame extraction + optical flow, both
    outside this path): `frames/*.png`, `flow/<direction>_<prev>_<this>.flo` and `.png`.  Decodes them, runs `vid_img_tensors` with
    the clip resident in HBM and writes `<size>/<pass>_<frame>.png` like the reference (:197, :294-297).  `frames`: the frame files
    in clip order (default: the sorted contents of `frames/`).  Not restated: skipping frames whose PNG already exists (resume),
    `--original_colors`, the ffmpeg encode of each finished scale (:302-304)."""
    import glob

    from PIL import Image

    _lib.require_gpu()
    if getattr(args, "original_colors", 0) == 1:
        raise NotImplementedError("maua_style_b200: --original_colors is host-side PIL work and not part of the device driver")
    dev = _device(args)
    work = str(args.output_dir) + "/" + _name(args.content) + "_" + "_".join(_name(p) for p in args.style)
    files = list(frames) if frames is not None else sorted(glob.glob(f"{work}/frames/*.png"))
    if len(files) < 2:
        raise FileNotFoundError(f"no frames under {work}/frames (run the reference's frame / flow preparation first)")
    names = [_name(f) for f in files]
    style_files = []
    for p in args.style:  # load.process_style_images (load.py:77-92): a directory stands for the images inside it
        if os.path.isdir(p):
            style_files += [f"{p}/{f}" for f in os.listdir(p) if os.path.splitext(f)[1].lower() in (".png", ".jpeg", ".jpg", ".tiff")]
        else:
            style_files.append(p)
    clip = [load_image(f, dev) for f in files]
    styles_big = [load_image(p, dev) for p in style_files]

    def flows(direction, i, j):
        stem = f"{work}/flow/{direction}_{names[i]}_{names[j]}"
        import numpy as np

        rel = np.asarray(Image.open(stem + ".png"), dtype=np.float32) / np.float32(255)  # T.ToTensor (load.py:217-218)
        rel = torch.from_numpy(rel if rel.ndim == 2 else rel[..., 0].copy())[None, None]
        return read_flo(stem + ".flo"), rel

    def on_frame(size, pass_n, f, u8):
        os.makedirs(f"{work}/{size}", exist_ok=True)
        Image.fromarray(u8.cpu().numpy(), mode="RGB").save(f"{work}/{size}/{pass_n}_{names[f]}.png")

    return vid_img_tensors(clip, styles_big, args, flows, on_frame=on_frame)
