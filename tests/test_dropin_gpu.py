"""The drop-in claim of INTEGRATION.md section 1, tested on the reference's own driver: the UNMODIFIED reference style.py
(style.py:22-73 img_img: load.preprocess -> F.interpolate -> optim.optimize per scale -> load.save_tensor_to_file), with
`loss` / `models` / `optim` swapped for maua_style_b200's through the sys.modules shim, against the PNGs the reference
wrote with its own modules (tests/golden/img_img_64_96.npz, made by tests/golden/make_golden_image.py).

The reference scripts are installed (copied unmodified, git-ignored) into baseline/_ref by __graft_entry__.build() where
/root/reference is mounted; the directory travels to the GPU box with the repo snapshot.
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from helpers import GOLDEN, ROOT, save_checkpoint

pytestmark = pytest.mark.gpu


def psnr_u8(a, b):
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


@pytest.fixture
def ref_b200():
    sys.path.insert(0, str(ROOT))
    from baseline import ref_loader

    if ref_loader.ref_dir() is None:
        pytest.skip("reference not installed: run __graft_entry__.build() where /root/reference is mounted")
    ref = ref_loader.import_reference("b200")
    yield ref, ref_loader
    ref_loader.unload()


def write_inputs(tmp_path):
    from PIL import Image

    z = np.load(GOLDEN / "img_img_64_96.npz", allow_pickle=False)
    for k in ("content", "style1", "style2"):
        Image.fromarray(z[k], mode="RGB").save(tmp_path / f"{k}.png")
    return z, json.loads(str(z["meta"]))


def run_reference_driver(ref, ref_loader, tmp_path, meta, **over):
    ckpt = tmp_path / "vgg19-random.pth"
    if not ckpt.exists():
        save_checkpoint(ckpt)
    args = ref_loader.reference_args(ref, tmp_path, ckpt, gpu="0", n_styles=2, optimizer=meta["optimizer"],
                                     style_blend_weights="3,1", image_sizes=",".join(map(str, meta["sizes"])),
                                     num_iters=",".join(map(str, meta["iters"])), init="content", **over)
    args.content = str(tmp_path / "content.png")
    args.style = [str(tmp_path / "style1.png"), str(tmp_path / "style2.png")]
    args.output = str(tmp_path / "out")
    args.match_histograms = False
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        ref.style.img_img(args)
    finally:
        os.chdir(cwd)
    from PIL import Image

    return [np.asarray(Image.open(tmp_path / f"out_{s}.png").convert("RGB")) for s in meta["sizes"]]


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_unmodified_reference_style_py_on_the_b200_modules(ref_b200, tmp_path, monkeypatch, precision):
    import maua_style_b200.optim as our_optim

    ref, ref_loader = ref_b200
    assert ref.style.optim is our_optim and ref.style.models.__name__ == "maua_style_b200.models"
    assert Path(ref.style.__file__).parent == ref.dir  # the driver itself is the reference's file
    monkeypatch.setenv("MAUA_PRECISION", precision)
    z, meta = write_inputs(tmp_path)
    outs = run_reference_driver(ref, ref_loader, tmp_path, meta)
    for size, got in zip(meta["sizes"], outs):
        want = z[f"out_{size}"]
        assert got.shape == want.shape
        p = psnr_u8(got, want)
        print(f"reference style.py + b200 modules ({precision}) {size}px vs reference PNG: PSNR {p:.1f} dB, "
              f"{int((got != want).sum())} of {got.size} bytes differ")
        assert p > (50.0 if precision == "fp32" else 38.0), (size, p)


def test_reference_driver_with_the_stock_scaling_presets(ref_b200, tmp_path, monkeypatch):
    """The stock config/scaling-img.json names models ("vgg19") instead of files and selects L-BFGS below 1456 px:
    optim.set_model_args overwrites args.model_file / optimizer with it, so the name must resolve through the model zoo."""
    ref, ref_loader = ref_b200
    z, meta = write_inputs(tmp_path)
    zoo = tmp_path / "modelzoo"
    zoo.mkdir()
    save_checkpoint(zoo / "vgg19.pth")
    monkeypatch.setenv("MAUA_MODELZOO", str(zoo))
    stock = Path(ref.dir) / "config" / "scaling-img.json"
    outs = run_reference_driver(ref, ref_loader, tmp_path, dict(meta, iters=[6, 4]), scaling_args=str(stock))
    assert [o.shape[:2] for o in outs] == [z["out_64"].shape[:2], z["out_96"].shape[:2]]
    # the mirror driver of this repo on the same inputs / presets writes the same PNGs (same kernels, same order)
    from maua_style_b200 import style as our_style

    ckpt = tmp_path / "vgg19-random.pth"
    args = ref_loader.reference_args(ref, tmp_path, ckpt, gpu="0", n_styles=2, style_blend_weights="3,1", image_sizes="64,96",
                                     num_iters="6,4", init="content", scaling_args=str(stock))
    args.content = str(tmp_path / "content.png")
    args.style = [str(tmp_path / "style1.png"), str(tmp_path / "style2.png")]
    args.output = str(tmp_path / "mirror")
    args.match_histograms = False
    args.style_scale = getattr(args, "style_scale", 1.0)
    our_style.img_img(args)
    from PIL import Image

    for size, got in zip([64, 96], outs):
        mirror = np.asarray(Image.open(tmp_path / f"mirror_{size}.png").convert("RGB"))
        p = psnr_u8(got, mirror)
        print(f"stock presets {size}px: reference driver vs mirror driver PSNR {p:.1f} dB")
        assert p > 45.0


def test_lbfgs_tolerance_grad_is_refused_not_ignored(tmp_path):
    from helpers import O, make_args
    from maua_style_b200 import optim

    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    args = make_args(ckpt, tmp_path, optimizer="lbfgs", lbfgs_tolerance_grad=1e-5)
    img = O.synthetic_image(64, 64, seed=1)
    with pytest.raises(NotImplementedError, match="lbfgs_tolerance_grad"):
        optim.optimize(img, [img], img.clone(), 2, args)
