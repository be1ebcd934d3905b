"""Golden vectors for the video driver around the per-frame optimisation (the `vid_img`
path), produced by the UNMODIFIED reference `style.vid_img` (style.py:145-296) on the CPU:

    python tests/golden/make_golden_video.py

  vid_img_3f_48_80.npz ... three 64x80 content frames, one style image, two scales (48 px, 80 px), two passes per scale
                           (forward, then backward over the reversed frame list), init=prev_warp, temporal_blend 0.5,
                           temporal_weight 50, Adam.  Stored: the input PNG bytes, the raw .flo flow fields and the
                           flow-reliability PNG bytes the driver read, and every PNG it wrote
                           (`out_<size>_<pass>_<frame>`).

The steps of the reference that are outside the scope of this repo are replaced by files written here:
`flow.get_flow_model` / `load.process_content_video` (frame extraction with ffmpeg + optical-flow estimation with the
CuPy networks) -- the frames, `.flo` fields and reliability maps they would have left in `<work_dir>/flow/` are seeded
synthetic data -- and the ffmpeg encode of each finished scale into an .mp4 (style.py:302-304), which is a no-op here.  Everything else (frame ordering, pass reversal, which PNG initialises / blends which frame, warp,
temporal targets, per-frame `optim.optimize`, PNG quantisation between passes) is the reference's own code.

  vid_img_loop_3f_48_80.npz  the same clip with --loop (style.py:181-183, :195-197): the frame list is rotated at a random start
                           every pass (python's global RNG, seeded with 7 here) and the first frames are styled a second
                           time from the frames the pass has just written (the `n > len(frames)` branches of :229-271).
  img_vid_9f_32_48.npz ... the UNMODIFIED reference `style.img_vid` (style.py:76-142): one 40x48 content image, one style
                           clip of 9 frames, two scales (32 px, 48 px) with frame windows of 3 and 2 frames, init=content
                           (content + seeded noise, blurred over time and space on the host), the 7-frame roll of pastiche
                           and style clips between scales (:134-135) and the temporal blur (:137-138), Adam.  Stored: the
                           content PNG bytes, the style clip, the initial pastiche video the driver built (its noise comes
                           from torch's global RNG) and the video tensor it handed to `load.save_tensor_to_file` after
                           every scale.  Replaced by stubs: `load.process_style_videos` (ffmpeg decode of the style clips)
                           and `load.save_tensor_to_file` (skvideo encode).

Adam, not L-BFGS, for the reason given in make_golden_image.py.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import import_reference, reference_args, save_checkpoint  # noqa: E402

FRAME_HW = (64, 80)
N_FRAMES = 3
SIZES, ITERS, PASSES = [48, 80], [8, 6], 2


def smooth_rgb(h, w, seed, drift=0.0):
    """A smooth seeded RGB image; `drift` shifts the low-frequency pattern so that consecutive frames look like motion."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(1, 3, max(h // 8, 2) + 2, max(w // 8, 2) + 2, generator=g)
    img = F.interpolate(low, size=(h + 16, w + 16), mode="bilinear", align_corners=False)[0]
    o = int(round(drift))
    img = img[:, 8:8 + h, 8 + o:8 + o + w]
    noise = torch.from_numpy(np.random.RandomState(seed + 100).randn(3, h, w).astype(np.float32)) * 5
    return (img * 255 + noise).clamp(0, 255).byte().permute(1, 2, 0).numpy()


def write_flo(path, flow):
    with open(path, "wb") as f:
        np.array([202021.25], dtype=np.float32).tofile(f)
        np.array([flow.shape[1]], dtype=np.int32).tofile(f)
        np.array([flow.shape[0]], dtype=np.int32).tofile(f)
        flow.astype(np.float32).tofile(f)


def make_img_vid(rconfig, rmodels, rload, rstyle):
    from PIL import Image

    import scipy.ndimage as ndi

    T, (H, W) = 9, (40, 48)
    sizes, iters, windows = [32, 48], [3, 2], "3,2"
    with tempfile.TemporaryDirectory(prefix="maua_golden_imgvid_") as tmp:
        workdir = Path(tmp)
        ckpt = workdir / "vgg19-random.pth"
        save_checkpoint(rmodels, ckpt)
        content_rgb = smooth_rgb(H, W, 1)
        Image.fromarray(content_rgb, mode="RGB").save(workdir / "content.png")
        mean = torch.tensor([103.939, 116.779, 123.68])[:, None, None]

        def pre(rgb):  # load.preprocess layout (load.py:21-32): BGR, 0-255, mean-subtracted
            return (torch.from_numpy(rgb.astype(np.float32)).permute(2, 0, 1)[[2, 1, 0]] - mean)[None]

        clip = torch.cat([pre(smooth_rgb(44, 52, 30, drift=1.0 * f)) for f in range(T)])
        args = reference_args(rconfig, workdir, ckpt, optimizer="adam", image_sizes=",".join(map(str, sizes)),
                              num_iters=",".join(map(str, iters)), init="content", transfer_type="img_vid",
                              gram_frame_window=windows, avg_frame_window=-1, num_frames=-1, temporal_blend=0.5, fps=24)
        args.content = str(workdir / "content.png")
        args.output = str(workdir / "out")
        args.match_histograms = False
        saved, inits = [], []
        rload.process_style_videos = lambda a: [clip.clone()]                          # ffmpeg decode: out of scope
        rload.save_tensor_to_file = lambda t, a, filename=None, **k: saved.append((filename, t.clone()))  # skvideo encode
        real_filter = ndi.gaussian_filter

        def spy(x, sigma, **k):  # the first call builds the initial pastiche video (style.py:99); record it
            y = real_filter(x, sigma, **k)
            if not inits:
                inits.append(np.array(y, dtype=np.float32))
            return y

        rstyle.ndi.gaussian_filter = spy
        torch.manual_seed(0)
        torch.set_flush_denormal(True)
        cwd = os.getcwd()
        os.chdir(workdir)
        try:
            rstyle.img_vid(args)
        finally:
            os.chdir(cwd)
            rstyle.ndi.gaussian_filter = real_filter
        out = {"content": content_rgb, "style_clip": clip.numpy().astype(np.float32), "init_video": inits[0]}
        assert len(saved) == len(sizes) + 1  # one per scale + the final file (style.py:141, :143)
        for s_, (fname, t) in zip(sizes, saved):
            assert fname.endswith(f"_{s_}")
            out[f"out_{s_}"] = t.numpy().astype(np.float32)
        out["meta"] = json.dumps(dict(sizes=sizes, iters=iters, windows=windows, T=T, init="content", optimizer="adam", temporal_blend=0.5,
                                      content_weight=args.content_weight, style_weight=args.style_weight, tv_weight=args.tv_weight,
                                      video_style_factor=args.video_style_factor))
        np.savez_compressed(HERE / "img_vid_9f_32_48.npz", **out)
        print("img_vid_9f_32_48.npz:", {k: getattr(v, "shape", None) for k, v in out.items() if not isinstance(v, str)})


def main():
    from PIL import Image

    rconfig, rloss, rmodels, roptim = import_reference()
    sys.modules["flow"].get_flow_model = lambda args: None  # optical-flow networks: out of scope, see the docstring

    class _NoEncode:  # style.py:302-304 encodes the PNGs of a scale into an .mp4 with ffmpeg: host I/O, not installed here
        def __getattr__(self, _):
            return lambda *a, **k: self

    sys.modules["ffmpeg"].input = lambda *a, **k: _NoEncode()
    import load as rload  # noqa
    import style as rstyle  # noqa

    if "--only-vid-img" not in sys.argv:
        make_img_vid(rconfig, rmodels, rload, rstyle)
    if "--only-img-vid" in sys.argv:
        return
    if "--only-loop" not in sys.argv:
        make_vid_img(rconfig, rmodels, rload, rstyle, loop=False, fname="vid_img_3f_48_80.npz", iters=ITERS)
    # --loop: random rotation of the frame list per pass (python's `random`, seeded here) + the first frames styled twice
    make_vid_img(rconfig, rmodels, rload, rstyle, loop=True, fname="vid_img_loop_3f_48_80.npz", iters=[4, 4])


def make_vid_img(rconfig, rmodels, rload, rstyle, loop, fname, iters):
    import random

    from PIL import Image

    ITERS = iters
    H, W = FRAME_HW
    with tempfile.TemporaryDirectory(prefix="maua_golden_vid_") as tmp:
        workdir = Path(tmp)
        ckpt = workdir / "vgg19-random.pth"
        save_checkpoint(rmodels, ckpt)
        (workdir / "in").mkdir()
        style_rgb = smooth_rgb(70, 90, 2)
        Image.fromarray(style_rgb, mode="RGB").save(workdir / "in" / "style.png")
        args = reference_args(rconfig, workdir, ckpt, optimizer="adam", image_sizes=",".join(map(str, SIZES)),
                              num_iters=",".join(map(str, ITERS)), init="prev_warp", transfer_type="vid_img",
                              temporal_weight=50.0, passes_per_scale=PASSES, loop=loop, temporal_blend=0.5)
        args.content = str(workdir / "in" / "clip.mp4")
        args.style = [str(workdir / "in" / "style.png")]
        args.match_histograms = False
        args.ffmpeg = {}
        work = Path(args.output_dir + "/clip_style")
        (work / "frames").mkdir(parents=True)
        (work / "flow").mkdir()
        out = {}
        frames = []
        for i in range(N_FRAMES):
            rgb = smooth_rgb(H, W, 1, drift=2.0 * i)
            p = work / "frames" / f"{i + 1:04d}.png"
            Image.fromarray(rgb, mode="RGB").save(p)
            frames.append(str(p))
            out[f"frame_{i}"] = rgb
        out["style"] = style_rgb
        rs = np.random.RandomState(17)
        fh, fw = H // 2, W // 2  # the flow files may have any resolution: flow_warp_map resizes the grid (load.py:211-213)
        for i in range(N_FRAMES):
            for direction, j in (("forward", (i + 1) % N_FRAMES), ("backward", (i - 1) % N_FRAMES)):
                sign = 1.0 if direction == "forward" else -1.0
                flow = np.zeros((fh, fw, 2), np.float32)
                flow[..., 0] = sign * 1.0 + rs.randn(fh, fw) * 0.6
                flow[..., 1] = rs.randn(fh, fw) * 0.6
                rel = (rs.rand(fh, fw) > 0.15).astype(np.uint8) * 255
                rel[rs.rand(fh, fw) > 0.9] = 128
                stem = f"{direction}_{i + 1:04d}_{j + 1:04d}"
                write_flo(work / "flow" / f"{stem}.flo", flow)
                Image.fromarray(rel, mode="L").save(work / "flow" / f"{stem}.png")
                out[f"flow_{direction}_{i}_{j}"] = flow
                out[f"rel_{direction}_{i}_{j}"] = rel
        rload.process_content_video = lambda model, a: list(frames)  # ffmpeg + flow estimation: out of scope
        torch.manual_seed(0)
        random.seed(7)  # style.py:182 draws the rotation of a --loop pass from python's global RNG
        torch.set_flush_denormal(True)
        cwd = os.getcwd()
        os.chdir(workdir)
        try:
            rstyle.vid_img(args)
        finally:
            os.chdir(cwd)
        for s in SIZES:
            for p in range(1, PASSES + 1):
                for i in range(N_FRAMES):
                    f = work / str(s) / f"{p}_{i + 1:04d}.png"
                    out[f"out_{s}_{p}_{i}"] = np.asarray(Image.open(f).convert("RGB"))
        out["meta"] = json.dumps(dict(sizes=SIZES, iters=ITERS, passes=PASSES, n_frames=N_FRAMES, init="prev_warp", optimizer="adam",
                                      loop=loop, random_seed=7, temporal_blend=0.5, temporal_weight=50.0, content_weight=args.content_weight,
                                      style_weight=args.style_weight, tv_weight=args.tv_weight))
        np.savez_compressed(HERE / fname, **out)
        print(fname + ":", {k: getattr(v, "shape", None) for k, v in out.items() if not isinstance(v, str)})


if __name__ == "__main__":
    main()
