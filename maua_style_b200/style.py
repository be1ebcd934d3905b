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

`img_img_tensors` / `stylize_frame` work on tensors (what benchmarks and the sharded runner call); `img_img(args)` is
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
    """One frame of vid_img (style.py:276-296) on the device: warp the previous frame's pastiche along the flow
    (`F.grid_sample(..., padding_mode="border")`, :279), resize the flow-reliability map (:283-286), capture the temporal
    target (:288), blend the initialisation (:290) and optimise (:292-294).  `net, losses` come from `models.load_model`
    once per scale (:176-177); the style targets are captured for the first frame only (optim.set_style_targets cache).
    Without `prev_pastiche` the frame starts from the content frame (style.py:225-226)."""
    dev = net.device
    with torch.cuda.device(dev):
        content_frame = content_frame.to(dev, torch.float32).contiguous()
        if prev_pastiche is None:
            pastiche = content_frame.clone()
        else:
            pastiche = prev_pastiche.to(dev, torch.float32).contiguous()
            if tuple(pastiche.shape[2:]) != tuple(content_frame.shape[2:]):
                pastiche = image_ops.interpolate(pastiche, size=tuple(content_frame.shape[2:]))  # style.py:241-243
            if flow_grid is not None:
                warp_image = image_ops.grid_sample(pastiche, flow_grid.to(dev))
                if reliable_flow is not None:
                    reliable_flow = image_ops.interpolate(reliable_flow.to(dev, torch.float32), size=tuple(pastiche.shape[2:]))
                optim.set_temporal_targets(net, warp_image, warp_weights=reliable_flow, args=args)
            if blend_image is not None:
                tb = float(getattr(args, "temporal_blend", 0.5))
                blend_image = blend_image.to(dev, torch.float32).contiguous()
                if tuple(blend_image.shape[2:]) != tuple(pastiche.shape[2:]):
                    blend_image = image_ops.interpolate(blend_image, size=tuple(pastiche.shape[2:]))  # style.py:253-255
                pastiche = image_ops.blend(blend_image, pastiche, 1.0 - tb, tb)
        return optim.optimize_device(content_frame, style_images, pastiche, num_iters, args, net, losses)


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
