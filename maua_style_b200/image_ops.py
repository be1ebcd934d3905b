"""Device-resident image operations either side of the optimisation loop, thin
wrappers over the C ABI (include/maua_b200.h, csrc/image_ops.cu).  Same call shapes as the torch functions the
reference uses, so `style.py`-like drivers read the same:

    interpolate(x, scale_factor=s | size=(h, w))   F.interpolate(..., mode="bilinear", align_corners=False)
                                                   reference style.py:38-41, :47-49, :57-66, :205-212, :241-255, :284-286
    grid_sample(x, grid)                           F.grid_sample(x, grid, padding_mode="border")   style.py:223, :279
    preprocess(img) / deprocess_u8(t)              load.py:21-32 / :47-52
    blend(x, y, a, b)                              style.py:290

Everything runs on the tensor's CUDA device on the current stream; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple, Union

import torch

from . import _lib


def _dev_f32(x: torch.Tensor, device=None) -> torch.Tensor:
    if not x.is_cuda:
        _lib.require_gpu()
        x = x.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    return x.to(torch.float32).contiguous()


def interp_out_size(n_in: int, scale_factor: float) -> int:
    """torch.nn.functional.interpolate's output extent for a scale factor: floor(float(in * scale))."""
    return int(math.floor(float(n_in * scale_factor)))


def interpolate(x: torch.Tensor, size: Optional[Sequence[int]] = None, scale_factor: Optional[float] = None,
                mode: str = "bilinear", align_corners: bool = False) -> torch.Tensor:
    """Bilinear resize of an [N,C,H,W] image, bit-identical to torch's CPU F.interpolate for fp32."""
    if mode != "bilinear" or align_corners:
        raise NotImplementedError("maua_style_b200.interpolate implements mode='bilinear', align_corners=False (all the "
                                  "reference uses)")
    if (size is None) == (scale_factor is None):
        raise ValueError("exactly one of size / scale_factor must be given")
    if x.dim() != 4:
        raise ValueError(f"expected an [N,C,H,W] tensor, got {tuple(x.shape)}")
    x = _dev_f32(x)
    n, c, hin, win = x.shape
    if scale_factor is not None:
        hout, wout = interp_out_size(hin, scale_factor), interp_out_size(win, scale_factor)
        sh = sw = float(scale_factor)
    else:
        hout, wout = int(size[0]), int(size[1])
        sh = sw = 0.0
    if hout <= 0 or wout <= 0:
        raise ValueError(f"interpolate: output size {hout}x{wout} is empty")
    out = torch.empty(n, c, hout, wout, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().maua_resize_bilinear(_lib.ptr(x), _lib.ptr(out), n * c, hin, win, hout, wout, C.c_double(sh),
                                                    C.c_double(sw), _lib.stream_ptr()), "maua_resize_bilinear")
    return out


def grid_sample(x: torch.Tensor, grid: torch.Tensor, padding_mode: str = "border") -> torch.Tensor:
    """F.grid_sample(x, grid, padding_mode="border") for x [1,C,H,W], grid [1,Ho,Wo,2]."""
    if padding_mode != "border":
        raise NotImplementedError("only padding_mode='border' (what the reference uses) is implemented")
    if x.dim() != 4 or grid.dim() != 4 or grid.shape[-1] != 2 or x.shape[0] != 1 or grid.shape[0] != 1:
        raise ValueError(f"expected x [1,C,H,W] and grid [1,Ho,Wo,2], got {tuple(x.shape)} and {tuple(grid.shape)}")
    x = _dev_f32(x)
    grid = _dev_f32(grid, x.device)
    _, c, hin, win = x.shape
    hout, wout = int(grid.shape[1]), int(grid.shape[2])
    out = torch.empty(1, c, hout, wout, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().maua_grid_sample_border(_lib.ptr(x), _lib.ptr(grid), _lib.ptr(out), c, hin, win, hout, wout,
                                                       _lib.stream_ptr()), "maua_grid_sample_border")
    return out


def preprocess(image: Union[torch.Tensor, "object"], device=None) -> torch.Tensor:
    """load.preprocess (load.py:21-32) for an image already in memory -> [1,3,H,W] BGR 0-255 mean-subtracted on the GPU.
    `image`: a uint8 [H,W,3] RGB tensor / numpy array / PIL image (its bytes travel over PCIe, 3 B per pixel), or a float
    [3,H,W] RGB tensor in [0,1] (ToTensor layout)."""
    _lib.require_gpu()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if not isinstance(image, torch.Tensor):
        import numpy as np

        arr = np.asarray(image.convert("RGB") if hasattr(image, "convert") else image)
        image = torch.from_numpy(np.ascontiguousarray(arr))
    if image.dtype == torch.uint8:
        if image.dim() != 3 or image.shape[2] != 3:
            raise ValueError(f"uint8 images must be [H,W,3] RGB, got {tuple(image.shape)}")
        src = image.to(dev).contiguous()
        h, w = int(src.shape[0]), int(src.shape[1])
        out = torch.empty(1, 3, h, w, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().maua_preprocess_u8(_lib.ptr(src), _lib.ptr(out), h, w, _lib.stream_ptr()), "maua_preprocess_u8")
        return out
    if image.dim() != 3 or image.shape[0] != 3:
        raise ValueError(f"float images must be [3,H,W] RGB in [0,1], got {tuple(image.shape)}")
    src = image.to(dev, torch.float32).contiguous()
    h, w = int(src.shape[1]), int(src.shape[2])
    out = torch.empty(1, 3, h, w, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().maua_preprocess_f32(_lib.ptr(src), _lib.ptr(out), h, w, _lib.stream_ptr()), "maua_preprocess_f32")
    return out


def deprocess_u8(t: torch.Tensor) -> torch.Tensor:
    """load.deprocess (load.py:47-52) down to the bytes of the PIL image: uint8 [H,W,3] RGB, still on the GPU."""
    if t.dim() == 4:
        if t.shape[0] != 1:
            raise ValueError("deprocess expects one image")
        t = t[0]
    if t.dim() != 3 or t.shape[0] != 3:
        raise ValueError(f"expected a [1,3,H,W] / [3,H,W] image, got {tuple(t.shape)}")
    t = _dev_f32(t)
    h, w = int(t.shape[1]), int(t.shape[2])
    out = torch.empty(h, w, 3, device=t.device, dtype=torch.uint8)
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().maua_deprocess_u8(_lib.ptr(t), _lib.ptr(out), h, w, _lib.stream_ptr()), "maua_deprocess_u8")
    return out


def deprocess(t: torch.Tensor):
    """load.deprocess: a PIL image (one 3 B/pixel device-to-host copy)."""
    from PIL import Image

    return Image.fromarray(deprocess_u8(t).cpu().numpy(), mode="RGB")


def blend(x: torch.Tensor, y: torch.Tensor, a: float, b: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a * x + b * y (style.py:290 with a = 1 - temporal_blend, b = temporal_blend)."""
    x = _dev_f32(x)
    y = _dev_f32(y, x.device)
    if x.shape != y.shape:
        raise ValueError(f"blend: shapes differ: {tuple(x.shape)} vs {tuple(y.shape)}")
    if out is None:
        out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().maua_blend(_lib.ptr(x), _lib.ptr(y), _lib.ptr(out), C.c_long(x.numel()), C.c_float(a), C.c_float(b),
                                          _lib.stream_ptr()), "maua_blend")
    return out


def flow_warp_grid(flow: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """The sampling grid of load.flow_warp_map (load.py:208-214) from an already normalised + smoothed flow field
    [h, w, 2] (load.py:201-206, host preprocessing): identity grid linspace(-1, 1) + flow, resized to `size`."""
    _lib.require_gpu()
    dev = flow.device if flow.is_cuda else torch.device("cuda", torch.cuda.current_device())
    h, w = int(flow.shape[0]), int(flow.shape[1])
    # load.py:208-210 verbatim arithmetic: float64 identity grid (numpy linspace) + flow, then one cast to fp32
    import numpy as np

    neutral = np.rollaxis(np.array(np.meshgrid(np.linspace(-1, 1, w), np.linspace(-1, 1, h))), 0, 3)
    warp = torch.from_numpy((neutral + flow.detach().cpu().numpy()).astype(np.float32))
    grid = interpolate(warp.permute(2, 0, 1).unsqueeze(0).to(dev), size=size)
    return grid.permute(0, 2, 3, 1).contiguous()
