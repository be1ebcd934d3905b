"""CPU oracle for the image-side steps either side of the maua-style optimisation loop.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as oracle/maua_oracle.py: only tests/, smoke() and
bench.py's CPU legs may import it).

Plain numpy fp32 restatements, operation by operation, of what the reference calls (citations relative to the
reference repo root, JCBrouwer/maua-style @ 316c552):

    resize_bilinear ...... F.interpolate(x, scale_factor=s | size=hw, mode="bilinear", align_corners=False)
                           style.py:38-41, :47-49, :57-66, :203-210, :242-254, :280-282; load.py:211-213
    grid_sample_border ... F.grid_sample(x, grid, padding_mode="border")            style.py:228, :276
    preprocess_u8 / _f32 . load.preprocess                                          load.py:21-32
    deprocess_u8 ......... load.deprocess + T.ToPILImage                             load.py:47-52
    blend ................ (1 - temporal_blend) * blend_image + temporal_blend * p   style.py:286
    match_histogram ...... utils.match_histogram (colour-statistics transfer)       utils.py:88-151
    img_img .............. the multi-resolution driver                               style.py:22-73
    flow_warp_map ........ .flo field -> sampling grid                               load.py:191-214
    vid_img .............. the per-frame video driver (scales x passes x frames)     style.py:145-300
    img_vid .............. the frame-window video driver (roll + temporal blur)      style.py:76-142

The arithmetic of interpolate / grid_sample lives in the third-party dependency PyTorch (pinned torch==1.8.1,
requirements.txt:1; ATen UpSampleBilinear2d / GridSampler, not in the reference tree); the published algorithm is
restated here: source index = fma(scale, dst+0.5, -0.5) clamped at 0 with scale = 1/scale_factor when a scale
factor was given, else in/out, value = fma(v0, w0, v1*w1) along w then h; grid_sample un-normalises with
fma(g+1, size/2, -0.5), clips to [0,size-1] and accumulates the four neighbours times the "opposite corner" areas
with fused multiply-adds.  The placement of the fused operations is that of the ATen CPU kernels (it was found by
matching torch 2.11's outputs bit for bit; with it the restatement is bit-exact on every golden).

Pinning: tests/golden/make_golden_image.py runs torch's own F.interpolate / F.grid_sample (the ops the reference
calls), the UNMODIFIED reference load.preprocess / load.deprocess and the UNMODIFIED reference style.img_img on
seeded inputs and commits the outputs (tests/golden/image_ops.npz, img_img_64_96.npz); tests/test_image_oracle.py
checks this file against them.  The video driver is pinned the same way: tests/golden/make_golden_video.py runs the
UNMODIFIED reference style.vid_img on seeded frames / .flo files (tests/golden/vid_img_3f_48_80.npz) and style.img_vid
on a seeded style clip (tests/golden/img_vid_9f_32_48.npz).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

BGR_MEAN = np.array([103.939, 116.779, 123.68], dtype=np.float32)  # load.py:30
f32 = np.float32


def interp_out_size(n_in: int, scale_factor: float) -> int:
    """torch.nn.functional.interpolate: floor(float(in * scale_factor))."""
    return int(math.floor(float(n_in * scale_factor)))


def _src_index(scale: np.float32, n_out: int, n_in: int):
    dst = np.arange(n_out, dtype=np.float32)
    # one fused multiply-add (the ATen build contracts scale*(dst+0.5)-0.5; float64 holds the product exactly)
    real = (np.float64(scale) * (dst + f32(0.5)).astype(np.float64) - 0.5).astype(np.float32)
    real = np.where(real < 0, f32(0), real).astype(np.float32)
    i0 = np.minimum(np.floor(real).astype(np.int64), n_in - 1)
    lam = np.clip(real - i0.astype(np.float32), f32(0), f32(1)).astype(np.float32)
    i1 = np.minimum(i0 + 1, n_in - 1)
    return i0, i1, (f32(1) - lam).astype(np.float32), lam


def resize_bilinear(x: np.ndarray, size: Optional[Tuple[int, int]] = None, scale_factor: Optional[float] = None) -> np.ndarray:
    """x: [..., H, W] float32.  Exactly one of size / scale_factor."""
    x = np.asarray(x, dtype=np.float32)
    hin, win = x.shape[-2:]
    if scale_factor is not None:
        hout, wout = interp_out_size(hin, scale_factor), interp_out_size(win, scale_factor)
        rh = rw = f32(1.0 / float(scale_factor))
    else:
        hout, wout = int(size[0]), int(size[1])
        rh, rw = f32(hin) / f32(hout), f32(win) / f32(wout)
    h0, h1, hl0, hl1 = _src_index(rh, hout, hin)
    w0, w1, wl0, wl1 = _src_index(rw, wout, win)

    def lerp(a, b, wa, wb):  # fma(a, wa, round(b * wb)) -- ATen's contraction of wa*a + wb*b
        return (a.astype(np.float64) * wa + (b * wb).astype(np.float32).astype(np.float64)).astype(np.float32)

    r0, r1 = x[..., h0, :], x[..., h1, :]
    if hout + wout <= 128:
        # ATen routes small outputs (_use_vectorized_kernel_cond_2d: out_h + out_w <= 128) to its vectorised kernel, which
        # multiplies the weights out first: w00*a + w01*b + w10*c + w11*d, contracted as fma(d,w11,fma(c,w10,fma(a,w00,b*w01)))
        a, b, c, d = r0[..., :, w0], r0[..., :, w1], r1[..., :, w0], r1[..., :, w1]
        w00, w01 = (hl0[:, None] * wl0[None, :]).astype(np.float32), (hl0[:, None] * wl1[None, :]).astype(np.float32)
        w10, w11 = (hl1[:, None] * wl0[None, :]).astype(np.float32), (hl1[:, None] * wl1[None, :]).astype(np.float32)
        acc = (a.astype(np.float64) * w00 + (b * w01).astype(np.float32).astype(np.float64)).astype(np.float32)
        acc = (c.astype(np.float64) * w10 + acc.astype(np.float64)).astype(np.float32)
        return (d.astype(np.float64) * w11 + acc.astype(np.float64)).astype(np.float32)
    top = lerp(r0[..., :, w0], r0[..., :, w1], wl0, wl1)
    bot = lerp(r1[..., :, w0], r1[..., :, w1], wl0, wl1)
    return lerp(top, bot, hl0[:, None], hl1[:, None])


def grid_sample_border(x: np.ndarray, grid: np.ndarray) -> np.ndarray:
    """x: [planes, H, W]; grid: [Ho, Wo, 2] normalised (x, y).  Bilinear, border padding, align_corners=False."""
    x = np.asarray(x, dtype=np.float32)
    g = np.asarray(grid, dtype=np.float32)
    hin, win = x.shape[-2:]
    # ATen CPU GridSampler: (g + 1) * (size / 2) - 0.5 as one fused multiply-add
    ix = ((g[..., 0] + f32(1)).astype(np.float64) * np.float64(f32(win) / f32(2)) - 0.5).astype(np.float32)
    iy = ((g[..., 1] + f32(1)).astype(np.float64) * np.float64(f32(hin) / f32(2)) - 0.5).astype(np.float32)
    ix = np.minimum(f32(win - 1), np.maximum(ix, f32(0)))
    iy = np.minimum(f32(hin - 1), np.maximum(iy, f32(0)))
    x0 = np.floor(ix).astype(np.int64)
    y0 = np.floor(iy).astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    w = ix - x0.astype(np.float32)
    e = f32(1) - w
    n = iy - y0.astype(np.float32)
    s = f32(1) - n
    nw, ne, sw, se = s * e, s * w, n * e, n * w
    bx1, by1 = x1 < win, y1 < hin
    x1c, y1c = np.minimum(x1, win - 1), np.minimum(y1, hin - 1)
    out = (x[:, y0, x0] * nw).astype(np.float32)
    for val, wt in ((np.where(bx1, x[:, y0, x1c], f32(0)), ne), (np.where(by1, x[:, y1c, x0], f32(0)), sw),
                    (np.where(bx1 & by1, x[:, y1c, x1c], f32(0)), se)):
        out = (val.astype(np.float64) * wt + out.astype(np.float64)).astype(np.float32)  # acc = fma(val, wt, acc)
    return out


def preprocess_u8(rgb_hwc: np.ndarray) -> np.ndarray:
    """uint8 [H,W,3] RGB -> float32 [1,3,H,W] BGR, 0-255, mean-subtracted (ToTensor's /255 and the reference's *255 kept)."""
    t = (rgb_hwc.astype(np.float32) / f32(255)).transpose(2, 0, 1) * f32(255)
    return (t[[2, 1, 0]] - BGR_MEAN[:, None, None]).astype(np.float32)[None]


def preprocess_f32(rgb_chw: np.ndarray) -> np.ndarray:
    t = np.asarray(rgb_chw, dtype=np.float32) * f32(255)
    return (t[[2, 1, 0]] - BGR_MEAN[:, None, None]).astype(np.float32)[None]


def deprocess_u8(bgr: np.ndarray) -> np.ndarray:
    """float32 [1,3,H,W] or [3,H,W] -> uint8 [H,W,3] RGB, the bytes of the PIL image load.deprocess returns."""
    t = np.asarray(bgr, dtype=np.float32).reshape(3, *bgr.shape[-2:])
    t = (t - (-BGR_MEAN)[:, None, None]).astype(np.float32)
    t = (t[[2, 1, 0]] / f32(255)).astype(np.float32)
    t = np.clip(t, f32(0), f32(1))
    return (t * f32(255)).astype(np.float32).astype(np.uint8).transpose(1, 2, 0)  # .byte() truncates


def blend(x: np.ndarray, y: np.ndarray, a: float, b: float) -> np.ndarray:
    return (f32(a) * np.asarray(x, np.float32) + f32(b) * np.asarray(y, np.float32)).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# utils.match_histogram (utils.py:88-151): colour-statistics transfer, SURVEY.md section 8f rank 3
# ---------------------------------------------------------------------------------------------------------------
def _sym_sqrt(c: np.ndarray) -> np.ndarray:
    """utils.py:124-127: Q = V sqrt(diag(w)) V^T of a symmetric matrix, NaNs of sqrt(negative eigenvalue) set to 0."""
    w, v = np.linalg.eigh(c)
    with np.errstate(invalid="ignore"):
        e = np.sqrt(w)
    e[np.isnan(e)] = 0
    return (v * e[None, :]) @ v.T


def channel_stats(img: np.ndarray, eps: float):
    """utils.get_histogram (utils.py:88-93) for one [1,3,H,W] image: per-channel mean over every pixel and the channel
    covariance h h^T / N + eps * I of the mean-centred values."""
    x = np.asarray(img, dtype=np.float64).reshape(3, -1)
    mu = x.mean(axis=1)
    h = x - mu[:, None]
    return mu, h @ h.T / h.shape[1] + eps * np.eye(3)


def match_histogram_affine(target: np.ndarray, sources: Sequence[np.ndarray], eps: float = 1e-2, noise_var: float = 1e-6):
    """(M, b) of the affine colour map `y = M x + b` that utils.match_histogram(target, sources, mode=True or "avg")
    applies to every pixel of a single-frame [1,3,H,W] target, in expectation over its noise: per source s,
    `Qs Qt^-1 (x - mu_t) + mu_s` with Q = sqrt of the channel covariance (+ eps I); the per-source results are averaged
    (utils.py:112-143).  The reference perturbs both images with 1e-3 * randn first (utils.py:120-121): that adds
    `noise_var` = 1e-6 to the covariance diagonals, which is kept here.  float64."""
    mu_t, ct = channel_stats(target, eps + noise_var)
    qt_inv = np.linalg.inv(_sym_sqrt(ct))
    m_bar, b_bar = np.zeros((3, 3)), np.zeros(3)
    for s in sources:
        mu_s, cs = channel_stats(s, eps + noise_var)
        m = _sym_sqrt(cs) @ qt_inv
        m_bar += m / len(sources)
        b_bar += (mu_s - m @ mu_t) / len(sources)
    return m_bar, b_bar


def match_histogram(target: np.ndarray, sources: Sequence[np.ndarray], eps: float = 1e-2, noise_var: float = 1e-6) -> np.ndarray:
    """utils.match_histogram without its unseeded per-pixel output perturbation (`1e-3 * M * randn`, the noise the
    reference adds to the image it transforms): `M x + b` in float64, returned as float32."""
    target = np.asarray(target, dtype=np.float32)
    m_bar, b_bar = match_histogram_affine(target, sources, eps, noise_var)
    x = target.astype(np.float64).reshape(3, -1)
    return (m_bar @ x + b_bar[:, None]).reshape(target.shape).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# style.py:22-73 img_img on tensors (no files, no histogram matching: match_histogram is a no-op on torch >= 2
# because th.symeig is gone -- SURVEY.md section 2 row 11 -- and goldens are generated with it disabled)
# ---------------------------------------------------------------------------------------------------------------
def img_img_plan(content_hw: Tuple[int, int], style_hws: Sequence[Tuple[int, int]], image_sizes: Sequence[int],
                 style_scale: float = 1.0):
    """Per scale: (content (h, w), content scale factor, [(style (h, w), style scale factor)]) -- style.py:36-50."""
    plan = []
    for size in image_sizes:
        cs = size / max(*content_hw)
        ch, cw = interp_out_size(content_hw[0], cs), interp_out_size(content_hw[1], cs)
        styles = []
        for sh, sw in style_hws:
            ss = math.sqrt((ch * cw) / (sw * sh)) * style_scale
            styles.append(((interp_out_size(sh, ss), interp_out_size(sw, ss)), ss))
        plan.append(((ch, cw), cs, styles))
    return plan


def img_img(content_big: np.ndarray, styles_big: List[np.ndarray], image_sizes: Sequence[int], num_iters: Sequence[int],
            optimize_fn, init: str = "content", style_scale: float = 1.0):
    """Drives `optimize_fn(content, styles, pastiche, num_iters) -> pastiche` over the scales.  Arrays are [1,3,H,W].
    Returns the list of per-scale results."""
    assert init == "content", "random init draws unseeded noise in the reference (style.py:55); goldens use init=content"
    outs, pastiche = [], None
    for size, iters in zip(image_sizes, num_iters):
        cs = size / max(*content_big.shape[-2:])
        content = resize_bilinear(content_big, scale_factor=cs)
        area = content.shape[2] * content.shape[3]
        styles = []
        for s in styles_big:
            ss = math.sqrt(area / (s.shape[3] * s.shape[2])) * style_scale
            styles.append(resize_bilinear(s, scale_factor=ss))
        src = content_big if pastiche is None else pastiche
        pastiche = resize_bilinear(src, size=content.shape[2:])
        pastiche = np.asarray(optimize_fn(content, styles, pastiche, iters), dtype=np.float32)
        outs.append(pastiche)
    return outs


# ---------------------------------------------------------------------------------------------------------------
# style.py:145-300 vid_img on arrays.  The PNG files the reference writes after every frame and reads back at the start
# of the next pass / scale are a dict of uint8 images here (the 8-bit quantisation between passes is part of the result).
# ---------------------------------------------------------------------------------------------------------------
def flow_warp_map(flow_raw: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """load.flow_warp_map (load.py:191-214) from the [h, w, 2] field of a .flo file: normalise by the field's own extent,
    Gaussian blur (sigma 5 along h and w), add the identity grid linspace(-1, 1), resize to `size`.  Returns [H, W, 2]."""
    import scipy.ndimage

    h, w = flow_raw.shape[:2]
    flow = np.array(flow_raw, dtype=np.float32)
    flow[:, :, 0] /= w
    flow[:, :, 1] /= h
    flow = scipy.ndimage.gaussian_filter(flow, [5, 5, 0])
    neutral = np.rollaxis(np.array(np.meshgrid(np.linspace(-1, 1, w), np.linspace(-1, 1, h))), 0, 3)
    warp = (neutral + flow).astype(np.float32)  # th.FloatTensor(float64 array): one rounding to fp32
    return resize_bilinear(warp.transpose(2, 0, 1)[None], size=tuple(size))[0].transpose(1, 2, 0)


def vid_img_schedule(n_frames: int, loop: bool = False):
    """style.py:192-194: the (previous frame, this frame) pairs of one pass over `order` (a list of frame indices).
    Without --loop every frame is `this frame` exactly once, the first one last (it follows the last frame)."""
    def pairs(order):
        return list(zip(order + order[: 11 if loop else 1], order[1:] + order[: 10 if loop else 1]))
    return pairs


def vid_img(frames_u8: Sequence[np.ndarray], styles_big: List[np.ndarray], image_sizes: Sequence[int], num_iters: Sequence[int],
            passes_per_scale: int, optimize_fn, flows, init: str = "prev_warp", temporal_blend: float = 0.5,
            style_scale: float = 1.0, chunks: Optional[Sequence[Sequence[int]]] = None):
    """Drives `optimize_fn(content, styles, pastiche, num_iters, temporal) -> pastiche` like style.vid_img does
    (`temporal` = None or (warped previous result, resized flow-reliability map), i.e. what optim.set_temporal_targets got).

    frames_u8: the decoded frames, uint8 [H,W,3] RGB; styles_big: preprocessed [1,3,h,w] arrays;
    flows(direction, prev_index, this_index) -> (raw .flo field [h,w,2], reliability PNG bytes uint8 [h,w]).
    Returns {(size, pass (1-based), frame index): uint8 [h,w,3]} -- the PNGs `<size>/<pass>_<frame>.png` (style.py:197).
    --loop (random rotation, style.py:183-185) and random init (unseeded, style.py:221-222) are not restated.

    `chunks` (lists of frame indices, one per GPU) restates what a job sharded over GPUs computes: whenever the owner of the
    frame changes, the chain of carried-over results breaks and the frame starts from the stored result of its predecessor in
    the previous pass / scale -- the reference's resume rule for frames whose PNG already exists (style.py:198-200, :232-271)."""
    owner = {f: r for r, c in enumerate(chunks or []) for f in c}
    assert init in ("prev_warp", "content"), init
    n = len(frames_u8)
    order = list(range(n))
    pairs = vid_img_schedule(n)
    big = [preprocess_u8(f) for f in frames_u8]
    H, W = big[0].shape[-2:]
    store = {}
    prev_size = None
    for size_n, (size, iters) in enumerate(zip(image_sizes, num_iters)):
        cs = size / max(H, W)
        area = cs ** 2 * H * W  # style.py:169: from the un-rounded scale, not from the resized frame
        styles = [resize_bilinear(s, scale_factor=math.sqrt(area / (s.shape[3] * s.shape[2])) * style_scale) for s in styles_big]
        for pass_n in range(passes_per_scale):
            pastiche = None
            direction = "forward" if pass_n % 2 == 0 else "backward"
            last_owner = None
            for prev_f, this_f in pairs(order):
                if owner and owner[this_f] != last_owner:
                    pastiche, last_owner = None, owner[this_f]
                content = [resize_bilinear(big[prev_f], scale_factor=cs), resize_bilinear(big[this_f], scale_factor=cs)]
                temporal = None
                if size_n == 0 and pass_n == 0:  # style.py:220-230
                    if init == "prev_warp":
                        if pastiche is None:
                            pastiche = content[0]
                        raw, _ = flows(direction, prev_f, this_f)
                        grid = flow_warp_map(raw, pastiche.shape[2:])
                        pastiche = grid_sample_border(pastiche[0], grid)[None]
                    else:
                        pastiche = content[1].copy()
                else:  # style.py:231-286
                    src = (prev_size, passes_per_scale) if pass_n == 0 else (size, pass_n)
                    if pastiche is None:
                        pastiche = preprocess_u8(store[src + (prev_f,)])
                        if pass_n == 0:
                            pastiche = resize_bilinear(pastiche, size=content[0].shape[2:])
                    blend_image = preprocess_u8(store[src + (this_f,)])
                    if pass_n == 0:
                        blend_image = resize_bilinear(blend_image, size=content[0].shape[2:])
                    raw, rel_u8 = flows(direction, prev_f, this_f)
                    grid = flow_warp_map(raw, pastiche.shape[2:])
                    warp_image = grid_sample_border(pastiche[0], grid)[None]
                    rel = (rel_u8.astype(np.float32) / f32(255))[None, None]  # T.ToTensor (load.py:217-218)
                    rel = resize_bilinear(rel, size=pastiche.shape[2:])
                    temporal = (warp_image, rel)
                    pastiche = blend(blend_image, pastiche, 1 - temporal_blend, temporal_blend)  # the UN-warped previous result
                pastiche = np.asarray(optimize_fn(content[1], styles, pastiche, iters // passes_per_scale, temporal), dtype=np.float32)
                store[(size, pass_n + 1, this_f)] = deprocess_u8(pastiche)
            order = list(reversed(order))  # style.py:300
        prev_size = size
    return store


# ---------------------------------------------------------------------------------------------------------------
# style.py:76-142 img_vid on arrays (the pastiche is a video [T,3,H,W] optimised in frame windows by optim.optimize)
# ---------------------------------------------------------------------------------------------------------------
def temporal_blur_wrap(video: np.ndarray, sigma: float) -> np.ndarray:
    """ndi.gaussian_filter(video, [sigma, 0, 0, 0], mode="wrap") (style.py:137-138) restated: a normalised Gaussian of radius
    int(4 sigma + 0.5) along the frame axis with periodic extension, accumulated in float64, result in the input's fp32."""
    r = int(4.0 * float(sigma) + 0.5)
    x = np.arange(-r, r + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    w = w / w.sum()
    v = np.asarray(video, dtype=np.float64)
    out = np.zeros_like(v)
    for k, wk in zip(x, w):
        out += wk * np.roll(v, -int(k), axis=0)  # out[t] += w[k] * v[(t + k) mod T]
    return out.astype(np.float32)


def img_vid(content_big: np.ndarray, style_clips_big: List[np.ndarray], init_video: np.ndarray, image_sizes: Sequence[int],
            num_iters: Sequence[int], frame_windows: Sequence[int], optimize_fn, temporal_blend: float = 0.5,
            style_scale: float = 1.0, roll: int = 7):
    """Drives `optimize_fn(content, style_clips, pastiche_video, num_iters, gram_frame_window) -> pastiche_video` over the
    scales like style.img_vid: resize content / style clips / pastiche (:114-130), optimise (:132), roll pastiche and style
    clips by 7 frames (:134-135), blur over time (:137-138).  `init_video` is the initial pastiche of :93-103 (its noise is
    drawn from the global RNG and blurred on the host; the golden stores it).  Returns the per-scale videos [T,3,h,w]."""
    H, W = content_big.shape[-2:]
    pastiche = np.asarray(init_video, dtype=np.float32)
    clips = [np.asarray(c, dtype=np.float32) for c in style_clips_big]
    outs = []
    for size, iters, gfw in zip(image_sizes, num_iters, frame_windows):
        content = resize_bilinear(content_big, scale_factor=size / max(H, W))
        area = content.shape[2] * content.shape[3]
        styles = [resize_bilinear(c, scale_factor=math.sqrt(area / (c.shape[3] * c.shape[2])) * style_scale) for c in clips]
        pastiche = resize_bilinear(pastiche, size=content.shape[2:])
        pastiche = np.asarray(optimize_fn(content, styles, pastiche, iters, int(gfw)), dtype=np.float32)
        pastiche = np.concatenate((pastiche[roll:], pastiche[:roll]))
        clips = [np.concatenate((c[roll:], c[:roll])) for c in clips]
        if temporal_blend > 0:
            pastiche = temporal_blur_wrap(pastiche, temporal_blend)
        outs.append(pastiche)
    return outs
