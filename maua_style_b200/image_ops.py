"""Device-resident image operations either side of the optimisation loop, thin
wrappers over the C ABI (include/maua_b200.h, csrc/image_ops.cu).  Same call shapes as the torch functions the
reference uses, so `style.py`-like drivers read the same:

    interpolate(x, scale_factor=s | size=(h, w))   F.interpolate(..., mode="bilinear", align_corners=False)
                                                   reference style.py:38-41, :47-49, :57-66, :203-210, :242-254, :280-282
    grid_sample(x, grid)                           F.grid_sample(x, grid, padding_mode="border")   style.py:228, :276
    preprocess(img) / deprocess_u8(t)              load.py:21-32 / :47-52
    blend(x, y, a, b)                              style.py:286
    match_histogram(target, sources, eps, mode)    utils.match_histogram   utils.py:88-151 (style.py:24, :67, :71)

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
    """a * x + b * y (style.py:286 with a = 1 - temporal_blend, b = temporal_blend)."""
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


# ---------------------------------------------------------------------------------------------------------------------
# utils.match_histogram (utils.py:88-151) -- SURVEY.md section 8f rank 3
# ---------------------------------------------------------------------------------------------------------------------
_REDUCE_WS = {}
MOMENT_NOISE_VAR = 1e-6  # variance of the `1e-3 * randn` perturbation of utils.py:120-121, as it enters the covariances


def _reduce_ws(dev: torch.device) -> torch.Tensor:
    """Zero-initialised scratch of the deterministic grid reductions, one per device (its counter re-arms itself)."""
    key = (dev.type, dev.index)
    if key not in _REDUCE_WS:
        _REDUCE_WS[key] = torch.zeros(_lib.load().maua_reduce_workspace_bytes(), dtype=torch.uint8, device=dev)
    return _REDUCE_WS[key]


def image_moments(img: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """{N, sum x_c, sum x_c x_d (upper triangle)} of one [1,3,H,W] / [3,H,W] image as 10 float64 values on the device: the
    sufficient statistics of utils.get_histogram (utils.py:88-93).  One pass over the image, deterministic."""
    if img.dim() == 4:
        if img.shape[0] != 1:
            raise NotImplementedError("image_moments takes one frame (the reference's B > 1 video paths are out of scope)")
        img = img[0]
    if img.dim() != 3 or img.shape[0] != 3:
        raise ValueError(f"expected a [1,3,H,W] / [3,H,W] image, got {tuple(img.shape)}")
    img = _dev_f32(img)
    if out is None:
        out = torch.empty(10, device=img.device, dtype=torch.float64)
    with torch.cuda.device(img.device):
        _lib.check(_lib.load().maua_image_moments(_lib.ptr(img), int(img.shape[1]), int(img.shape[2]), _lib.ptr(out),
                                                  _lib.ptr(_reduce_ws(img.device)), _lib.stream_ptr()), "maua_image_moments")
    return out


def match_histogram(target_tensor: torch.Tensor, source_tensor, eps: float = 1e-2, mode="avg",
                    source_moments: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """utils.match_histogram(target_tensor, source_tensor, eps, mode) for single-frame images, on the device: the target's
    channel means / covariances are mapped onto each source's and the results averaged (utils.py:112-143).  A falsy
    `mode` returns the target unchanged (utils.py:97-98); for one frame "avg" and the random-frame mode coincide.

    Two passes over the target (moments, affine map) + one per source; the 3x3 eigen-decompositions run in one device
    thread, so nothing returns to the host.  `source_moments` ([n,10] float64 from `image_moments`) skips the source
    passes when the same style images are matched against repeatedly (every scale of style.py:67-71).  The statistics
    include the variance of the reference's `1e-3 * randn` input perturbation (utils.py:120-121); its unseeded per-pixel
    effect on the output is not reproduced."""
    if not mode:
        return target_tensor
    if target_tensor.dim() != 4 or target_tensor.shape[0] != 1 or target_tensor.shape[1] != 3:
        raise NotImplementedError(f"match_histogram takes one [1,3,H,W] frame, got {tuple(target_tensor.shape)}")
    target = _dev_f32(target_tensor)
    dev = target.device
    if source_moments is None:
        sources = source_tensor if isinstance(source_tensor, (list, tuple)) else [source_tensor]
        source_moments = torch.empty(len(sources), 10, device=dev, dtype=torch.float64)
        for i, s in enumerate(sources):
            image_moments(_dev_f32(s, dev), out=source_moments[i])
    source_moments = source_moments.to(dev, torch.float64).contiguous().view(-1, 10)
    h, w = int(target.shape[2]), int(target.shape[3])
    if out is None:
        out = torch.empty_like(target)
    affine = torch.empty(12, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        lib = _lib.load()
        tm = image_moments(target)
        _lib.check(lib.maua_hist_match_coefs(_lib.ptr(tm), _lib.ptr(source_moments), int(source_moments.shape[0]),
                                             C.c_double(eps + MOMENT_NOISE_VAR), _lib.ptr(affine), _lib.stream_ptr()),
                   "maua_hist_match_coefs")
        _lib.check(lib.maua_color_affine(_lib.ptr(target), _lib.ptr(out), h, w, _lib.ptr(affine), _lib.stream_ptr()),
                   "maua_color_affine")
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
