"""Golden vectors for the image-side steps around the optimisation loop, produced by
the code the reference itself runs:

    python tests/golden/make_golden_image.py

  image_ops.npz ........ torch's own F.interpolate(mode="bilinear", align_corners=False) / F.grid_sample(padding_mode=
                         "border") on seeded inputs (the third-party ops the reference calls at style.py:38-66, :223,
                         :279), and the UNMODIFIED reference load.preprocess / load.deprocess / load.flow_warp_map
                         (load.py:21-52, :191-214) on a seeded PNG / .flo file
  img_img_64_96.npz .... the UNMODIFIED reference style.img_img (style.py:22-73) end to end on seeded PNGs: two scales
                         (64 px, 96 px), two blended styles of different sizes, init=content, Adam; the PNGs it
                         writes are stored as uint8 arrays.  (Adam, not L-BFGS: from init=content the first L-BFGS step is
                         1/|g|_1 ~ 5e-6 long, so y = g1 - g0 is round-off noise and two fp32 implementations of the
                         same algorithm already disagree at 29-43 dB -- not a usable fixture.)

The reference cannot travel to the GPU box, so the outputs are committed; tests/test_image_oracle.py pins
oracle/image_oracle.py to them on the CPU and tests/test_image_gpu.py checks the CUDA path against both.
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
from make_golden import REF, import_reference, reference_args, save_checkpoint  # noqa: E402

RESIZE_CASES = [  # (h, w, scale_factor, size, seed)
    (64, 96, 0.5, None, 11), (90, 122, 1.37, None, 12), (61, 45, None, (128, 77), 13), (128, 128, 724 / 1024, None, 14),
    (100, 60, 0.333, None, 15), (40, 50, 1.21, None, 16), (30, 30, None, (61, 47), 17), (96, 72, 2.0, None, 18),
    (181, 90, 1448 / 1024, None, 19), (7, 9, None, (64, 64), 20), (200, 200, None, (1, 1), 21),
]


def seeded(shape, seed, lo=-120.0, hi=140.0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * (hi - lo) + lo


def make_image_ops(rload, workdir: Path):
    from PIL import Image

    out = {"resize_cases": json.dumps(RESIZE_CASES)}
    for i, (h, w, sf, size, seed) in enumerate(RESIZE_CASES):
        x = seeded((1, 3, h, w), seed)
        if sf is not None:
            y = F.interpolate(x, scale_factor=sf, mode="bilinear", align_corners=False)
        else:
            y = F.interpolate(x, size=size, mode="bilinear", align_corners=False)
        out[f"resize_{i}"] = y.numpy()
    # grid_sample with coordinates beyond [-1, 1] (border clamp) and exactly on the borders
    x = seeded((1, 3, 50, 70), 31)
    g = seeded((1, 40, 60, 2), 32, -1.25, 1.25)
    g[0, 0, :4, 0] = torch.tensor([-1.0, 1.0, 0.0, 1.0 - 1.0 / 70])
    g[0, 0, :4, 1] = torch.tensor([1.0, -1.0, 0.0, -1.0 + 1.0 / 50])
    out["grid_x"], out["grid_g"] = x.numpy(), g.numpy()
    out["grid_y"] = F.grid_sample(x, g, padding_mode="border").numpy()
    # load.preprocess on a seeded PNG
    rs = np.random.RandomState(5)
    rgb = rs.randint(0, 256, size=(37, 53, 3)).astype(np.uint8)
    rgb[0, :8, :] = np.arange(8, dtype=np.uint8)[:, None] * 36 + 3  # includes 255 and small values
    rgb[1, 0] = (0, 128, 255)
    png = workdir / "pre.png"
    Image.fromarray(rgb, mode="RGB").save(png)
    out["pre_rgb"] = rgb
    out["pre_out"] = rload.preprocess(str(png)).numpy()
    # load.deprocess on values inside and outside the displayable range
    t = seeded((1, 3, 29, 41), 6, -160.0, 190.0)
    t[0, :, 0, 0] = torch.tensor([-103.939, -116.779, -123.68])              # exactly black
    t[0, :, 0, 1] = torch.tensor([255 - 103.939, 255 - 116.779, 255 - 123.68])  # exactly white
    out["de_in"] = t.numpy()
    out["de_out"] = np.asarray(rload.deprocess(t.clone()))
    # style.py:290 blend
    a, b = seeded((1, 3, 20, 30), 7), seeded((1, 3, 20, 30), 8)
    out["blend_a"], out["blend_b"] = a.numpy(), b.numpy()
    out["blend_out"] = ((1 - 0.35) * a + 0.35 * b).numpy()
    # load.flow_warp_map on a seeded .flo file
    fh, fw = 20, 30
    flow = (np.random.RandomState(9).randn(fh, fw, 2) * 3.0).astype(np.float32)
    flo = workdir / "f.flo"
    with open(flo, "wb") as f:
        np.array([202021.25], dtype=np.float32).tofile(f)
        np.array([fw], dtype=np.int32).tofile(f)
        np.array([fh], dtype=np.int32).tofile(f)
        flow.tofile(f)
    import scipy.ndimage

    smooth = flow.copy()  # load.py:203-206: the host-side part (normalise + Gaussian blur), input of flow_warp_grid
    smooth[:, :, 0] /= fw
    smooth[:, :, 1] /= fh
    smooth = scipy.ndimage.gaussian_filter(smooth, [5, 5, 0])
    out["flow_smooth"] = smooth.astype(np.float32)
    out["flow_grid"] = rload.flow_warp_map(str(flo), (41, 57)).numpy()
    np.savez_compressed(HERE / "image_ops.npz", **out)
    print("image_ops.npz:", {k: getattr(v, "shape", None) for k, v in out.items() if not isinstance(v, str)})


def make_img_img(rconfig, rmodels, rstyle, workdir: Path, ckpt: Path):
    from PIL import Image

    rs = np.random.RandomState(21)

    def smooth_png(path, h, w, seed):
        g = torch.Generator().manual_seed(seed)
        low = torch.rand(1, 3, max(h // 8, 2), max(w // 8, 2), generator=g)
        img = F.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)[0]
        img = (img * 255 + torch.from_numpy(rs.randn(3, h, w).astype(np.float32)) * 6).clamp(0, 255).byte()
        arr = img.permute(1, 2, 0).numpy()
        Image.fromarray(arr, mode="RGB").save(path)
        return arr

    content = smooth_png(workdir / "content.png", 80, 112, 1)
    s1 = smooth_png(workdir / "style1.png", 70, 90, 2)
    s2 = smooth_png(workdir / "style2.png", 100, 60, 3)
    sizes, iters = [64, 96], [12, 8]
    args = reference_args(rconfig, workdir, ckpt, n_styles=2, optimizer="adam", style_blend_weights="3,1",
                          image_sizes=",".join(map(str, sizes)), num_iters=",".join(map(str, iters)), init="content")
    args.content = str(workdir / "content.png")
    args.style = [str(workdir / "style1.png"), str(workdir / "style2.png")]
    args.output = str(workdir / "out")
    args.match_histograms = False
    torch.manual_seed(0)
    torch.set_flush_denormal(True)
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        rstyle.img_img(args)
    finally:
        os.chdir(cwd)
    out = {"meta": json.dumps(dict(sizes=sizes, iters=iters, blend=[float(x) for x in args.style_blend_weights], optimizer="adam",
                                   content_weight=args.content_weight, style_weight=args.style_weight, tv_weight=args.tv_weight)),
           "content": content, "style1": s1, "style2": s2}
    for s in sizes:
        out[f"out_{s}"] = np.asarray(Image.open(workdir / f"out_{s}.png").convert("RGB"))
    np.savez_compressed(HERE / "img_img_64_96.npz", **out)
    print("img_img_64_96.npz:", {k: getattr(v, "shape", None) for k, v in out.items() if not isinstance(v, str)})


def main():
    rconfig, rloss, rmodels, roptim = import_reference()
    import load as rload  # noqa  (reference load.py; needs the skvideo stub installed by import_reference)
    import style as rstyle  # noqa
    with tempfile.TemporaryDirectory(prefix="maua_golden_img_") as tmp:
        workdir = Path(tmp)
        ckpt = workdir / "vgg19-random.pth"
        save_checkpoint(rmodels, ckpt)
        make_image_ops(rload, workdir)
        make_img_img(rconfig, rmodels, rstyle, workdir, ckpt)


if __name__ == "__main__":
    main()
