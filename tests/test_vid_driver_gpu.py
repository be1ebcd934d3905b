"""The vid_img driver (style.py:145-300: scales x passes x frames around `optim.optimize`) on
the device, against the PNGs the UNMODIFIED reference wrote on the CPU (tests/golden/vid_img_3f_48_80.npz, made by
tests/golden/make_golden_video.py; oracle/image_oracle.vid_img is pinned to the same file in tests/test_image_oracle.py).

Two routes to the same frames:
  * `maua_style_b200.style.vid_img_tensors` -- the whole schedule resident in HBM;
  * the reference's own `style.vid_img`, unmodified, on top of maua_style_b200.{loss,models,optim} (INTEGRATION.md
    section 1), reading / writing its PNG, .flo and reliability files.
Optical-flow estimation and ffmpeg frame extraction / encoding are outside the path: the flow files are inputs.
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from helpers import GOLDEN, ROOT, make_args, save_checkpoint

pytestmark = pytest.mark.gpu

# 2 scales x 2 passes x 3 frames of 4+1 / 3+1 Adam evaluations each, every frame starting from the previous results: in
# exact arithmetic only summation order separates the device from the reference's CPU run; with TF32 operands the rounding
# flips of section 2 of DESIGN.md accumulate along the chain, hence a PSNR bound (same rule as the img_img driver test)
MIN_PSNR = {"fp32": 48.0, "tf32": 40.0}  # measured on a B200: 52.7 / 46.3 dB (profiles/r05_vid_driver.txt)


def psnr_u8(a, b):
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def load_golden():
    z = np.load(GOLDEN / "vid_img_3f_48_80.npz", allow_pickle=False)
    return z, json.loads(str(z["meta"]))


def compare(z, meta, get, what):
    worst = 99.0
    for size in meta["sizes"]:
        for p in range(1, meta["passes"] + 1):
            for f in range(meta["n_frames"]):
                got, ref = get(size, p, f), z[f"out_{size}_{p}_{f}"]
                assert got.shape == ref.shape, (size, p, f, got.shape, ref.shape)
                worst = min(worst, psnr_u8(got, ref))
    print(f"{what}: worst PSNR over {len(meta['sizes']) * meta['passes'] * meta['n_frames']} frames {worst:.1f} dB")
    return worst


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_vid_img_tensors_matches_reference_pngs(tmp_path, monkeypatch, precision):
    from maua_style_b200 import image_ops, models, style

    monkeypatch.setenv("MAUA_PRECISION", precision)
    models.clear_model_cache()
    z, meta = load_golden()
    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    a = make_args(ckpt, tmp_path, transfer_type="vid_img", optimizer=meta["optimizer"], content_weight=meta["content_weight"],
                  style_weight=meta["style_weight"], tv_weight=meta["tv_weight"], temporal_weight=meta["temporal_weight"],
                  image_sizes=list(meta["sizes"]), num_iters=list(meta["iters"]), passes_per_scale=meta["passes"], init=meta["init"],
                  temporal_blend=meta["temporal_blend"], loop=False, style_scale=1.0, match_histograms=False)
    frames = [image_ops.preprocess(torch.from_numpy(z[f"frame_{i}"])) for i in range(meta["n_frames"])]
    flo_dir = tmp_path / "flow"
    flo_dir.mkdir()

    def flows(direction, i, j):
        raw = z[f"flow_{direction}_{i}_{j}"]
        path = flo_dir / f"{direction}_{i}_{j}.flo"
        with open(path, "wb") as f:  # Middlebury .flo, as the reference's flow stage leaves it (load.py:221-232)
            np.array([202021.25], dtype=np.float32).tofile(f)
            np.array([raw.shape[1]], dtype=np.int32).tofile(f)
            np.array([raw.shape[0]], dtype=np.int32).tofile(f)
            raw.astype(np.float32).tofile(f)
        rel = torch.from_numpy(z[f"rel_{direction}_{i}_{j}"].astype(np.float32) / np.float32(255))[None, None]  # T.ToTensor
        return style.read_flo(str(path)), rel

    seen = []
    store = style.vid_img_tensors(frames, [image_ops.preprocess(torch.from_numpy(z["style"]))], a, flows,
                                  on_frame=lambda size, p, f, u8: seen.append((size, p, f)))
    # frame order: forward passes end on frame 0 (it follows the last frame), backward passes run over the reversed list
    assert [s[2] for s in seen[:6]] == [1, 2, 0, 1, 0, 2]
    assert len(store) == len(seen) == len(meta["sizes"]) * meta["passes"] * meta["n_frames"]
    assert all(v.is_cuda and v.dtype == torch.uint8 for v in store.values())
    worst = compare(z, meta, lambda s, p, f: store[(s, p, f)].cpu().numpy(), f"vid_img_tensors ({precision}) vs reference PNGs")
    assert worst > MIN_PSNR[precision], worst
    # one network per scale for all passes and frames; the style targets of a scale are captured once (:176-177 + target cache)
    assert models.cache_stats["misses"] >= 1


def test_vid_img_file_level_entry(tmp_path):
    """style.vid_img(args) on the device: frames / .flo / reliability PNGs in the reference's directory layout in, the reference's
    `<size>/<pass>_<frame>.png` out, against the PNGs the unmodified reference wrote."""
    from PIL import Image

    from maua_style_b200 import style

    z, meta = load_golden()
    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    work = tmp_path / "out" / "clip_style"
    (work / "frames").mkdir(parents=True)
    (work / "flow").mkdir()
    Image.fromarray(z["style"], mode="RGB").save(tmp_path / "style.png")
    n = meta["n_frames"]
    for i in range(n):
        Image.fromarray(z[f"frame_{i}"], mode="RGB").save(work / "frames" / f"{i + 1:04d}.png")
        for d, j in (("forward", (i + 1) % n), ("backward", (i - 1) % n)):
            stem = work / "flow" / f"{d}_{i + 1:04d}_{j + 1:04d}"
            raw = z[f"flow_{d}_{i}_{j}"]
            with open(f"{stem}.flo", "wb") as f:
                np.array([202021.25], dtype=np.float32).tofile(f)
                np.array([raw.shape[1]], dtype=np.int32).tofile(f)
                np.array([raw.shape[0]], dtype=np.int32).tofile(f)
                raw.astype(np.float32).tofile(f)
            Image.fromarray(z[f"rel_{d}_{i}_{j}"], mode="L").save(f"{stem}.png")
    a = make_args(ckpt, tmp_path, transfer_type="vid_img", optimizer=meta["optimizer"], content_weight=meta["content_weight"],
                  style_weight=meta["style_weight"], tv_weight=meta["tv_weight"], temporal_weight=meta["temporal_weight"],
                  image_sizes=list(meta["sizes"]), num_iters=list(meta["iters"]), passes_per_scale=meta["passes"], init=meta["init"],
                  temporal_blend=meta["temporal_blend"], loop=False, style_scale=1.0, match_histograms=False, original_colors=0,
                  output_dir=str(tmp_path / "out"), content=str(tmp_path / "clip.mp4"), style=[str(tmp_path / "style.png")])
    store = style.vid_img(a)
    get = lambda s, p, f: np.asarray(Image.open(work / str(s) / f"{p}_{f + 1:04d}.png").convert("RGB"))
    assert all(np.array_equal(get(s, p, f), v.cpu().numpy()) for (s, p, f), v in store.items())
    assert compare(z, meta, get, "style.vid_img (files) vs reference PNGs") > MIN_PSNR["tf32"]


def test_sharded_video_with_one_rank_is_the_plain_driver(tmp_path, monkeypatch):
    """shard.stylize_video on one GPU owns every frame and exchanges with nobody: bit-identical to style.vid_img_tensors.  (Two
    ranks: tests/test_shard.py on gloo, against the oracle's chunked restatement.)"""
    from maua_style_b200 import image_ops, models, shard, style

    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    z, meta = load_golden()
    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)

    def args():
        return make_args(ckpt, tmp_path, transfer_type="vid_img", optimizer=meta["optimizer"], content_weight=meta["content_weight"],
                         style_weight=meta["style_weight"], tv_weight=meta["tv_weight"], temporal_weight=meta["temporal_weight"],
                         image_sizes=list(meta["sizes"])[:1], num_iters=list(meta["iters"])[:1], passes_per_scale=meta["passes"],
                         init=meta["init"], temporal_blend=meta["temporal_blend"], loop=False, style_scale=1.0, match_histograms=False)

    frames = [image_ops.preprocess(torch.from_numpy(z[f"frame_{i}"])) for i in range(meta["n_frames"])]
    styles = [image_ops.preprocess(torch.from_numpy(z["style"]))]

    def flows(direction, i, j):
        raw = z[f"flow_{direction}_{i}_{j}"].astype(np.float32).copy()
        raw[:, :, 0] /= raw.shape[1]
        raw[:, :, 1] /= raw.shape[0]
        import scipy.ndimage

        rel = torch.from_numpy(z[f"rel_{direction}_{i}_{j}"].astype(np.float32) / np.float32(255))[None, None]
        return torch.from_numpy(scipy.ndimage.gaussian_filter(raw, [5, 5, 0])), rel

    plain = style.vid_img_tensors(frames, styles, args(), flows)
    sharded = shard.stylize_video(frames, styles, args(), flows, shard.RankInfo(0, 1, 0))
    assert sorted(plain) == sorted(sharded) and len(plain) == meta["passes"] * meta["n_frames"]
    for k in plain:
        assert torch.equal(plain[k], sharded[k]), k
    # a schedule that draws random numbers cannot be sharded: refused, not silently different per rank
    a = args()
    a.loop = True
    with pytest.raises(NotImplementedError):
        style.vid_img_tensors(frames, styles, a, flows, owned=[0, 1])


def test_vid_img_tensors_with_histogram_matching(tmp_path):
    """style.py:209 / :294 with match_histograms on: content frames and every result take the colour statistics of the first
    style image (on torch >= 2 the reference's own call is a silent no-op; this is what it did under its pinned torch 1.8.1)."""
    from helpers import O
    from maua_style_b200 import image_ops, style

    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    a = make_args(ckpt, tmp_path, transfer_type="vid_img", image_sizes=[48], num_iters=[4], passes_per_scale=2, init="content",
                  temporal_blend=0.5, loop=False, style_scale=1.0, match_histograms=True)
    frames = [O.synthetic_image(48, 64, seed=60 + i, smooth=True) for i in range(3)]
    sty = (O.synthetic_image(56, 60, seed=2, smooth=True) * 0.5 + 20).cuda()
    flow = torch.zeros(12, 16, 2)
    rel = torch.ones(1, 1, 12, 16)
    store = style.vid_img_tensors(frames, [sty], a, lambda d, i, j: (flow, rel))
    ms = image_ops.image_moments(sty).cpu().numpy()
    assert len(store) == 6
    for key, u8 in store.items():
        mo = image_ops.image_moments(image_ops.preprocess(u8)).cpu().numpy()
        # 8-bit truncation shifts the mean by about half a grey level and adds 1/12 to the variance
        assert np.allclose(mo[1:4] / mo[0], ms[1:4] / ms[0], atol=1.0), key
        assert np.allclose(mo[[4, 7, 9]] / mo[0] - (mo[1:4] / mo[0]) ** 2, ms[[4, 7, 9]] / ms[0] - (ms[1:4] / ms[0]) ** 2, rtol=3e-2), key


@pytest.fixture
def ref_b200():
    sys.path.insert(0, str(ROOT))
    from baseline import ref_loader

    if ref_loader.ref_dir() is None:
        pytest.skip("reference not installed: run __graft_entry__.build() where /root/reference is mounted")
    ref = ref_loader.import_reference("b200")
    yield ref, ref_loader
    ref_loader.unload()


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_unmodified_reference_vid_img_on_the_b200_modules(ref_b200, tmp_path, monkeypatch, precision):
    """The reference's own video driver with its loss / models / optim modules swapped for this package's."""
    from PIL import Image

    import maua_style_b200.optim as our_optim

    ref, ref_loader = ref_b200
    assert ref.style.optim is our_optim
    monkeypatch.setenv("MAUA_PRECISION", precision)
    z, meta = load_golden()
    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    (tmp_path / "in").mkdir()
    Image.fromarray(z["style"], mode="RGB").save(tmp_path / "in" / "style.png")
    args = ref_loader.reference_args(ref, tmp_path, ckpt, gpu="0", optimizer=meta["optimizer"],
                                     image_sizes=",".join(map(str, meta["sizes"])), num_iters=",".join(map(str, meta["iters"])),
                                     init=meta["init"], transfer_type="vid_img", temporal_weight=meta["temporal_weight"],
                                     passes_per_scale=meta["passes"], loop=False, temporal_blend=meta["temporal_blend"])
    args.content = str(tmp_path / "in" / "clip.mp4")
    args.style = [str(tmp_path / "in" / "style.png")]
    args.match_histograms = False
    args.ffmpeg = {}
    work = Path(args.output_dir + "/clip_style")
    (work / "frames").mkdir(parents=True)
    (work / "flow").mkdir()
    n = meta["n_frames"]
    frames = []
    for i in range(n):
        p = work / "frames" / f"{i + 1:04d}.png"
        Image.fromarray(z[f"frame_{i}"], mode="RGB").save(p)
        frames.append(str(p))
    for i in range(n):
        for direction, j in (("forward", (i + 1) % n), ("backward", (i - 1) % n)):
            stem = work / "flow" / f"{direction}_{i + 1:04d}_{j + 1:04d}"
            raw = z[f"flow_{direction}_{i}_{j}"]
            with open(f"{stem}.flo", "wb") as f:
                np.array([202021.25], dtype=np.float32).tofile(f)
                np.array([raw.shape[1]], dtype=np.int32).tofile(f)
                np.array([raw.shape[0]], dtype=np.int32).tofile(f)
                raw.astype(np.float32).tofile(f)
            Image.fromarray(z[f"rel_{direction}_{i}_{j}"], mode="L").save(f"{stem}.png")

    # the two stages of the reference outside this path (make_golden_video.py stubs the same two): flow networks + ffmpeg
    class _NoEncode:
        def __getattr__(self, _):
            return lambda *a, **k: self

    monkeypatch.setattr(sys.modules["flow"], "get_flow_model", lambda a: None, raising=False)
    monkeypatch.setattr(sys.modules["ffmpeg"], "input", lambda *a, **k: _NoEncode(), raising=False)
    monkeypatch.setattr(ref.load, "process_content_video", lambda model, a: list(frames))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        ref.style.vid_img(args)
    finally:
        os.chdir(cwd)
    get = lambda s, p, f: np.asarray(Image.open(work / str(s) / f"{p}_{f + 1:04d}.png").convert("RGB"))
    worst = compare(z, meta, get, f"reference style.vid_img + b200 modules ({precision}) vs reference PNGs")
    assert worst > MIN_PSNR[precision], worst


# ---- img_vid: the frame-window video driver (style.py:76-142) ----
# Adam, 2 scales, 4 + 6 windows of 3 / 2 frames; measured on a B200 (profiles/r05_vid_driver.txt): exact mode 168.6 / 94.4 dB
# (the CPU oracle's own restatement: 170.8 / 94.4 dB), TF32 operands 56.6 / 51.0 dB
IMG_VID_MIN_PSNR = {"fp32": 80.0, "tf32": 44.0}


def load_img_vid_golden():
    z = np.load(GOLDEN / "img_vid_9f_32_48.npz", allow_pickle=False)
    return z, json.loads(str(z["meta"]))


def video_psnr(a, b):
    mse = float(((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2).mean())
    return 199.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_img_vid_tensors_matches_reference_videos(tmp_path, monkeypatch, precision):
    from maua_style_b200 import image_ops, models, style

    monkeypatch.setenv("MAUA_PRECISION", precision)
    models.clear_model_cache()
    z, meta = load_img_vid_golden()
    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    a = make_args(ckpt, tmp_path, transfer_type="img_vid", optimizer=meta["optimizer"], content_weight=meta["content_weight"],
                  style_weight=meta["style_weight"], tv_weight=meta["tv_weight"], video_style_factor=meta["video_style_factor"],
                  image_sizes=list(meta["sizes"]), num_iters=list(meta["iters"]), gram_frame_window=meta["windows"],
                  avg_frame_window=-1, num_frames=-1, temporal_blend=meta["temporal_blend"], init="content", style_scale=1.0,
                  match_histograms=False)
    content = image_ops.preprocess(torch.from_numpy(z["content"]))
    outs = style.img_vid_tensors(content, [torch.from_numpy(z["style_clip"])], a, init_video=torch.from_numpy(z["init_video"]))
    for size, out in zip(meta["sizes"], outs):
        ref = z[f"out_{size}"]
        assert out.is_cuda and tuple(out.shape) == ref.shape
        p = video_psnr(out.cpu().numpy(), ref)
        print(f"img_vid_tensors ({precision}) {size}px vs reference video: PSNR {p:.1f} dB")
        assert p > IMG_VID_MIN_PSNR[precision], (size, p)
    # the temporal blur on the device is scipy's (style.py:137-138)
    import scipy.ndimage as ndi

    v = torch.randn(9, 3, 20, 24, generator=torch.Generator().manual_seed(3)) * 60
    got = style.temporal_blur(v.cuda(), 0.5).cpu().numpy()
    assert float(np.abs(got - ndi.gaussian_filter(v.numpy(), [0.5, 0, 0, 0], mode="wrap")).max()) <= 4e-5


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_unmodified_reference_img_vid_on_the_b200_modules(ref_b200, tmp_path, monkeypatch, precision):
    """The reference's own style.img_vid with loss / models / optim swapped; its video decode / encode (ffmpeg, skvideo) are
    stubbed exactly like in tests/golden/make_golden_video.py."""
    from PIL import Image

    ref, ref_loader = ref_b200
    monkeypatch.setenv("MAUA_PRECISION", precision)
    z, meta = load_img_vid_golden()
    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    Image.fromarray(z["content"], mode="RGB").save(tmp_path / "content.png")
    args = ref_loader.reference_args(ref, tmp_path, ckpt, gpu="0", optimizer=meta["optimizer"],
                                     image_sizes=",".join(map(str, meta["sizes"])), num_iters=",".join(map(str, meta["iters"])),
                                     init="content", transfer_type="img_vid", gram_frame_window=meta["windows"],
                                     avg_frame_window=-1, num_frames=-1, temporal_blend=meta["temporal_blend"], fps=24)
    args.content = str(tmp_path / "content.png")
    args.output = str(tmp_path / "out")
    args.match_histograms = False
    saved = []
    clip = torch.from_numpy(z["style_clip"])
    monkeypatch.setattr(ref.load, "process_style_videos", lambda a: [clip.clone()])
    monkeypatch.setattr(ref.load, "save_tensor_to_file", lambda t, a, filename=None, **k: saved.append((filename, t.clone())))
    torch.manual_seed(0)  # the driver draws the initial pastiche from the global RNG (style.py:95-99), as in the golden run
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        ref.style.img_vid(args)
    finally:
        os.chdir(cwd)
    assert len(saved) == len(meta["sizes"]) + 1
    for size, (fname, out) in zip(meta["sizes"], saved):
        p = video_psnr(out.numpy(), z[f"out_{size}"])
        print(f"reference style.img_vid + b200 modules ({precision}) {size}px vs reference video: PSNR {p:.1f} dB")
        assert p > IMG_VID_MIN_PSNR[precision], (size, p)
