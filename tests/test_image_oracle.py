"""oracle/image_oracle.py against the golden vectors produced by torch's own ops and the unmodified reference
(tests/golden/make_golden_image.py).  CPU only.  Bit-exact: these are gathers, fused multiply-adds and casts."""
import json

import numpy as np
import pytest
import torch

from helpers import GOLDEN, O, save_checkpoint
from oracle import image_oracle as I


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "image_ops.npz", allow_pickle=False)


def seeded(shape, seed, lo=-120.0, hi=140.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * (hi - lo) + lo).numpy()


def resize_cases(gold):
    return [(i, *c) for i, c in enumerate(json.loads(str(gold["resize_cases"])))]


def test_resize_bilinear_bit_exact(gold):
    for i, h, w, sf, size, seed in resize_cases(gold):
        x = seeded((1, 3, h, w), seed)
        y = I.resize_bilinear(x, size=size, scale_factor=sf)
        ref = gold[f"resize_{i}"]
        assert y.shape == ref.shape, (i, y.shape, ref.shape)
        assert np.array_equal(y, ref), (i, h, w, sf, size, float(np.abs(y - ref).max()))


def test_interp_out_size_matches_torch():
    import torch.nn.functional as F

    for n, s in [(1024, 724 / 1024), (724, 1448 / 724), (122, 0.37), (90, 1.9999), (256, 2.0), (181, 4 / 3)]:
        assert I.interp_out_size(n, s) == F.interpolate(torch.zeros(1, 1, n, 8), scale_factor=s, mode="bilinear",
                                                         align_corners=False).shape[2]


def test_grid_sample_border_bit_exact(gold):
    y = I.grid_sample_border(gold["grid_x"][0], gold["grid_g"][0])
    assert np.array_equal(y, gold["grid_y"][0]), float(np.abs(y - gold["grid_y"][0]).max())


def test_preprocess_deprocess_bit_exact(gold):
    assert np.array_equal(I.preprocess_u8(gold["pre_rgb"]), gold["pre_out"])
    assert np.array_equal(I.deprocess_u8(gold["de_in"]), gold["de_out"])
    # ToTensor()-layout float input gives the same image as the uint8 route
    f = (gold["pre_rgb"].astype(np.float32) / np.float32(255)).transpose(2, 0, 1)
    assert np.array_equal(I.preprocess_f32(f), gold["pre_out"])


def test_blend_and_flow_grid(gold):
    assert np.array_equal(I.blend(gold["blend_a"], gold["blend_b"], 1 - 0.35, 0.35), gold["blend_out"])
    fs = gold["flow_smooth"]
    h, w = fs.shape[:2]
    neutral = np.rollaxis(np.array(np.meshgrid(np.linspace(-1, 1, w), np.linspace(-1, 1, h))), 0, 3)
    warp = (neutral + fs).astype(np.float32)
    grid = I.resize_bilinear(warp.transpose(2, 0, 1)[None], size=gold["flow_grid"].shape[1:3])[0].transpose(1, 2, 0)
    assert np.array_equal(grid, gold["flow_grid"][0])


def test_img_img_driver_matches_reference_pngs(tmp_path):
    """style.py:22-73 end to end: the oracle's driver + the oracle's Adam must reproduce the PNGs the unmodified
    reference wrote (fp32 CPU arithmetic on both sides; summation order inside torch ops is the only freedom)."""
    z = np.load(GOLDEN / "img_img_64_96.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    params = O.he_init_vgg19(0)
    cfg = O.StyleConfig(content_weight=meta["content_weight"], style_weight=meta["style_weight"], tv_weight=meta["tv_weight"],
                        optimizer=meta["optimizer"], style_blend_weights=meta["blend"])
    torch.set_flush_denormal(True)

    def optimize_fn(content, styles, pastiche, iters):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        return O.optimize(t(content), [t(s) for s in styles], t(pastiche), iters, cfg, params).detach().numpy()

    outs = I.img_img(I.preprocess_u8(z["content"]), [I.preprocess_u8(z["style1"]), I.preprocess_u8(z["style2"])],
                     meta["sizes"], meta["iters"], optimize_fn)
    for size, out in zip(meta["sizes"], outs):
        ref = z[f"out_{size}"]
        got = I.deprocess_u8(out)
        assert got.shape == ref.shape
        diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
        mse = float((diff.astype(np.float64) ** 2).mean())
        psnr = float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
        assert psnr > 45.0, (size, psnr, int(diff.max()))


def load_vid_golden():
    z = np.load(GOLDEN / "vid_img_3f_48_80.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    frames = [z[f"frame_{i}"] for i in range(meta["n_frames"])]
    flows = lambda d, a, b: (z[f"flow_{d}_{a}_{b}"], z[f"rel_{d}_{a}_{b}"])
    return z, meta, frames, flows


def psnr_u8(a, b):
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def test_vid_img_driver_matches_reference_pngs():
    """style.py:145-300 end to end: frame order and pass reversal, which stored frame initialises / blends which frame,
    the prev_warp start of the very first pass, warp + flow-reliability temporal targets, 8-bit quantisation between
    passes -- every PNG the unmodified reference wrote (2 scales x 2 passes x 3 frames) must be reproduced."""
    z, meta, frames, flows = load_vid_golden()
    params = O.he_init_vgg19(0)
    cfg = O.StyleConfig(content_weight=meta["content_weight"], style_weight=meta["style_weight"], tv_weight=meta["tv_weight"],
                        temporal_weight=meta["temporal_weight"], optimizer=meta["optimizer"])
    torch.set_flush_denormal(True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))

    def optimize_fn(content, styles, pastiche, iters, temporal):
        tmp = None if temporal is None else (t(temporal[0]), t(temporal[1]))
        return O.optimize(t(content), [t(s) for s in styles], t(pastiche), iters, cfg, params, temporal=tmp).detach().numpy()

    store = I.vid_img(frames, [I.preprocess_u8(z["style"])], meta["sizes"], meta["iters"], meta["passes"], optimize_fn, flows,
                      init=meta["init"], temporal_blend=meta["temporal_blend"])
    assert len(store) == len(meta["sizes"]) * meta["passes"] * meta["n_frames"]
    worst = 99.0
    for (size, p, f), got in store.items():
        ref = z[f"out_{size}_{p}_{f}"]
        assert got.shape == ref.shape, (size, p, f, got.shape, ref.shape)
        worst = min(worst, psnr_u8(got, ref))
        assert psnr_u8(got, ref) > 45.0, (size, p, f, psnr_u8(got, ref))
    print(f"vid_img oracle vs reference PNGs: worst PSNR {worst:.1f} dB")


def test_vid_img_schedule_visits_every_frame_once_per_pass():
    pairs = I.vid_img_schedule(4)
    assert pairs([0, 1, 2, 3]) == [(0, 1), (1, 2), (2, 3), (3, 0)]
    assert pairs([3, 2, 1, 0]) == [(3, 2), (2, 1), (1, 0), (0, 3)]
    # --loop styles the first frames a second time so that the end meets the start (zip stops at the shorter list)
    looped = I.vid_img_schedule(4, loop=True)([0, 1, 2, 3])
    assert looped == [(0, 1), (1, 2), (2, 3), (3, 0), (0, 1), (1, 2), (2, 3)]
    assert len(I.vid_img_schedule(24, loop=True)(list(range(24)))) == 24 + 9  # frames[1:] + frames[:10]


def test_img_vid_driver_matches_reference_videos():
    """style.py:76-142 end to end: per-scale frame windows, the 7-frame roll of pastiche and style clips, the temporal blur --
    the video tensors the unmodified reference handed to load.save_tensor_to_file after each scale must be reproduced."""
    z = np.load(GOLDEN / "img_vid_9f_32_48.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    params = O.he_init_vgg19(0)
    cfg = O.StyleConfig(content_weight=meta["content_weight"], style_weight=meta["style_weight"], tv_weight=meta["tv_weight"],
                        video_style_factor=meta["video_style_factor"], optimizer=meta["optimizer"])
    torch.set_flush_denormal(True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))

    def optimize_fn(content, styles, pastiche, iters, gfw):
        return O.optimize_windows(t(content), [t(s) for s in styles], t(pastiche), iters, cfg, params, gfw=gfw).detach().numpy()

    outs = I.img_vid(I.preprocess_u8(z["content"]), [z["style_clip"]], z["init_video"], meta["sizes"], meta["iters"],
                     [int(w) for w in meta["windows"].split(",")], optimize_fn, temporal_blend=meta["temporal_blend"])
    for size, out in zip(meta["sizes"], outs):
        ref = z[f"out_{size}"]
        assert out.shape == ref.shape, (size, out.shape, ref.shape)
        p = O.psnr(t(out), t(ref))
        print(f"img_vid oracle {size}px vs reference video: PSNR {p:.1f} dB")
        assert p > 80.0, (size, p)  # measured 170.8 / 94.4 dB


def test_temporal_blur_matches_scipy():
    import scipy.ndimage as ndi

    rs = np.random.RandomState(4)
    for T, sigma in [(9, 0.5), (5, 1.0), (3, 2.0), (12, 0.3)]:  # incl. a radius longer than the clip (periodic extension)
        v = (rs.randn(T, 3, 6, 7) * 50).astype(np.float32)
        ref = ndi.gaussian_filter(v, [sigma, 0, 0, 0], mode="wrap")
        got = I.temporal_blur_wrap(v, sigma)
        assert float(np.abs(got - ref).max()) <= 4e-5, (T, sigma, float(np.abs(got - ref).max()))  # measured: bit-equal


def hist_cases():
    g = np.load(GOLDEN / "hist_match.npz", allow_pickle=False)
    for name, thw, shws, mode, seed in json.loads(str(g["cases"])):
        yield name, g[f"{name}_target"], [g[f"{name}_source{i}"] for i in range(len(shws))], g[f"{name}_out"]


def check_hist_match(name, y, target, sources, ref):
    """`y` against the output of the reference's utils.match_histogram: the reference transforms `target + 1e-3 * randn`
    (utils.py:120), so the residual must be exactly that noise pushed through the colour map -- zero-mean with a
    per-channel standard deviation of 1e-3 * |row of M| (averaged over the sources, each with its own draw)."""
    eps = 1e-2 + 1e-6
    mu_t, ct = I.channel_stats(target, eps)
    qt_inv = np.linalg.inv(I._sym_sqrt(ct))
    ms = [I._sym_sqrt(I.channel_stats(s, eps)[1]) @ qt_inv for s in sources]
    sigma = 1e-3 * np.sqrt(sum(np.linalg.norm(m, axis=1) ** 2 for m in ms)) / len(ms)
    d = (np.asarray(y, dtype=np.float64) - ref)[0]
    n = d[0].size
    for c in range(3):
        rms = float(np.sqrt((d[c] ** 2).mean()))
        assert 0.85 * sigma[c] < rms < 1.15 * sigma[c], (name, c, rms, sigma[c])
        assert float(np.abs(d[c]).max()) < 5.5 * sigma[c], (name, c, float(np.abs(d[c]).max()), sigma[c])
        assert abs(float(d[c].mean())) < 5 * sigma[c] / np.sqrt(n) + 2e-5, (name, c, float(d[c].mean()))


def test_match_histogram_against_reference_outputs():
    for name, target, sources, ref in hist_cases():
        check_hist_match(name, I.match_histogram(target, sources), target, sources, ref)


def test_match_histogram_properties():
    """Size-independent properties: the matched image has the sources' mean and (single source) covariance up to eps;
    matching an image to itself is the identity; two identical sources equal one."""
    for name, target, sources, _ in hist_cases():
        y = I.match_histogram(target, sources)
        mu_y, c_y = I.channel_stats(y, 0.0)
        mu_s = np.mean([I.channel_stats(s, 0.0)[0] for s in sources], axis=0)
        assert np.allclose(mu_y, mu_s, atol=2e-4), (name, mu_y, mu_s)
        if len(sources) == 1 and name != "flat_target":
            c_s = I.channel_stats(sources[0], 0.0)[1]
            assert np.allclose(c_y, c_s, rtol=2e-3, atol=0.05), (name, np.abs(c_y - c_s).max())
        assert np.allclose(I.match_histogram(target, [target]), target, atol=2e-4)
        assert np.allclose(I.match_histogram(target, [sources[0], sources[0]]), I.match_histogram(target, [sources[0]]), atol=1e-4)


# ---------------------------------------------------------------------------------------------------------------------
# Randomised pinning: the numpy oracle against the torch CPU ops the reference calls (style.py:38-66 F.interpolate,
# style.py:279 F.grid_sample) on shapes and scale factors the fixtures do not contain.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(12))
def test_resize_bilinear_random_shapes_bit_exact_vs_torch(seed):
    import torch
    import torch.nn.functional as F

    from oracle import image_oracle as IO

    rs = np.random.RandomState(seed)
    h, w = int(rs.randint(3, 97)), int(rs.randint(3, 97))
    x = (rs.rand(1, 3, h, w).astype(np.float32) * 255.0 - 110.0)
    if seed % 2 == 0:
        sf = float(rs.uniform(0.3, 2.7))
        ref = F.interpolate(torch.from_numpy(x), scale_factor=sf, mode="bilinear", align_corners=False).numpy()
        out = IO.resize_bilinear(x, scale_factor=sf)
    else:
        size = (int(rs.randint(2, 130)), int(rs.randint(2, 130)))
        ref = F.interpolate(torch.from_numpy(x), size=size, mode="bilinear", align_corners=False).numpy()
        out = IO.resize_bilinear(x, size=size)
    assert out.shape == ref.shape
    if torch.get_num_threads() == 1 and not np.array_equal(out, ref):
        # with a single thread ATen's CPU kernel takes another loop and places its fused multiply-adds differently (results move by
        # up to 2 ulp); the committed fixtures (test_resize_bilinear_bit_exact) pin the multi-threaded form, which is also the
        # device kernel's
        assert float(np.abs(out - ref).max()) <= 4 * np.spacing(np.float32(np.abs(ref).max())), float(np.abs(out - ref).max())
        return
    assert np.array_equal(out, ref), float(np.abs(out - ref).max())


@pytest.mark.parametrize("seed", range(6))
def test_grid_sample_border_random_bit_exact_vs_torch(seed):
    import torch
    import torch.nn.functional as F

    from oracle import image_oracle as IO

    rs = np.random.RandomState(100 + seed)
    h, w = int(rs.randint(4, 70)), int(rs.randint(4, 70))
    x = (rs.rand(1, 3, h, w).astype(np.float32) * 255.0 - 110.0)
    grid = (rs.rand(1, h, w, 2).astype(np.float32) * 2.4 - 1.2)  # some samples fall outside: padding_mode="border"
    ref = F.grid_sample(torch.from_numpy(x), torch.from_numpy(grid), padding_mode="border", align_corners=False).numpy()
    out = IO.grid_sample_border(x[0], grid[0])
    assert np.array_equal(out, ref[0]), float(np.abs(out - ref).max())
