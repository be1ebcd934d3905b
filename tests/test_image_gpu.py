"""CUDA image-side operations and the device-resident drivers (maua_style_b200/image_ops.py, style.py; SURVEY.md section
8f ranks 1-2) against the golden vectors of torch's own ops / the unmodified reference and against the CPU oracle.
Gathers, fused multiply-adds and casts: bit-exact.  Drivers (they contain the TF32 optimisation loop): PSNR-bounded."""
import json
import math

import numpy as np
import pytest
import torch

from helpers import GOLDEN, O, make_args, rel, save_checkpoint
from oracle import image_oracle as I

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "image_ops.npz", allow_pickle=False)


def seeded(shape, seed, lo=-120.0, hi=140.0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * (hi - lo) + lo


def psnr_u8(a, b):
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return float("inf") if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)


def test_resize_bilinear_bit_exact_vs_torch_golden(gold):
    from maua_style_b200 import image_ops

    for i, (h, w, sf, size, seed) in enumerate(json.loads(str(gold["resize_cases"]))):
        x = seeded((1, 3, h, w), seed).cuda()
        y = image_ops.interpolate(x, size=size, scale_factor=sf).cpu().numpy()
        ref = gold[f"resize_{i}"]
        assert y.shape == ref.shape
        assert np.array_equal(y, ref), (i, h, w, sf, size, float(np.abs(y - ref).max()))


@pytest.mark.parametrize("h,w,sf,size", [(1024, 1024, 1448 / 1024, None), (724, 724, None, (1024, 1024)), (1448, 1448, 0.5, None),
                                         (512, 384, None, (2048, 1536)), (2048, 2048, 1024 / 2048, None)])
def test_resize_bilinear_full_size_vs_oracle(h, w, sf, size):
    """BASELINE.json's sizes (the 256 -> ... -> 1448 schedule and the 2048 case) against the numpy restatement."""
    from maua_style_b200 import image_ops

    x = seeded((1, 3, h, w), 7)
    y = image_ops.interpolate(x.cuda(), size=size, scale_factor=sf).cpu().numpy()
    ref = I.resize_bilinear(x.numpy(), size=size, scale_factor=sf)
    assert np.array_equal(y, ref), float(np.abs(y - ref).max())


def test_resize_properties_at_full_size():
    from maua_style_b200 import image_ops

    x = seeded((1, 3, 1024, 1024), 3).cuda()
    assert torch.equal(image_ops.interpolate(x, size=(1024, 1024)), x)            # identity
    c = torch.full((1, 3, 1024, 768), 37.25, device="cuda")
    assert torch.equal(image_ops.interpolate(c, scale_factor=1448 / 1024), torch.full((1, 3, 1448, 1086), 37.25, device="cuda"))
    up = image_ops.interpolate(x, scale_factor=2.0)                                  # convexity: no overshoot
    assert float(up.max()) <= float(x.max()) and float(up.min()) >= float(x.min())
    with pytest.raises(NotImplementedError):
        image_ops.interpolate(x, scale_factor=2.0, mode="nearest")
    with pytest.raises(ValueError):
        image_ops.interpolate(x)


def test_grid_sample_border_bit_exact(gold):
    from maua_style_b200 import image_ops

    y = image_ops.grid_sample(torch.from_numpy(gold["grid_x"]).cuda(), torch.from_numpy(gold["grid_g"]).cuda())
    assert np.array_equal(y.cpu().numpy(), gold["grid_y"])
    # identity grid at full size reproduces the image up to the interpolation weights being exactly 0 / 1
    h = w = 1024
    x = seeded((1, 3, h, w), 5).cuda()
    xs = (torch.arange(w, dtype=torch.float64) * 2 + 1) / w - 1
    ys = (torch.arange(h, dtype=torch.float64) * 2 + 1) / h - 1
    grid = torch.stack([xs[None, :].expand(h, w), ys[:, None].expand(h, w)], -1)[None].float().cuda()
    y = image_ops.grid_sample(x, grid)
    ref = I.grid_sample_border(x[0].cpu().numpy(), grid[0].cpu().numpy())
    assert np.array_equal(y[0].cpu().numpy(), ref)
    assert float((y - x).abs().max()) < 0.05


def test_preprocess_deprocess_bit_exact(gold):
    from maua_style_b200 import image_ops

    assert np.array_equal(image_ops.preprocess(torch.from_numpy(gold["pre_rgb"])).cpu().numpy(), gold["pre_out"])
    f = (torch.from_numpy(gold["pre_rgb"]).float() / 255).permute(2, 0, 1).contiguous()
    assert np.array_equal(image_ops.preprocess(f).cpu().numpy(), gold["pre_out"])
    assert np.array_equal(image_ops.deprocess_u8(torch.from_numpy(gold["de_in"]).cuda()).cpu().numpy(), gold["de_out"])
    # full size: a 2048^2 8-bit image survives preprocess -> deprocess unchanged wherever (v/255)*255 truncates back to v
    rgb = torch.from_numpy(np.random.RandomState(1).randint(0, 256, size=(2048, 2048, 3)).astype(np.uint8))
    back = image_ops.deprocess_u8(image_ops.preprocess(rgb)).cpu()
    assert np.array_equal(back.numpy(), I.deprocess_u8(I.preprocess_u8(rgb.numpy())))
    assert int((back.int() - rgb.int()).abs().max()) <= 1


def test_blend_and_flow_grid(gold):
    from maua_style_b200 import image_ops

    out = image_ops.blend(torch.from_numpy(gold["blend_a"]).cuda(), torch.from_numpy(gold["blend_b"]).cuda(), 1 - 0.35, 0.35)
    assert np.array_equal(out.cpu().numpy(), gold["blend_out"])
    grid = image_ops.flow_warp_grid(torch.from_numpy(gold["flow_smooth"]), tuple(gold["flow_grid"].shape[1:3]))
    assert np.array_equal(grid.cpu().numpy(), gold["flow_grid"])


def _img_args(tmp_path, meta, **over):
    ckpt = tmp_path / "vgg19-random.pth"
    if not ckpt.exists():
        save_checkpoint(ckpt)
    a = make_args(ckpt, tmp_path, optimizer=meta["optimizer"], content_weight=meta["content_weight"],
                  style_weight=meta["style_weight"], tv_weight=meta["tv_weight"], style_blend_weights=list(meta["blend"]),
                  image_sizes=list(meta["sizes"]), num_iters=list(meta["iters"]), init="content", style_scale=1.0, **over)
    return a


def test_img_img_driver_matches_reference_pngs(tmp_path):
    """style.py:22-73 end to end on the device against the PNGs written by the unmodified reference."""
    from maua_style_b200 import image_ops, models, style

    z = np.load(GOLDEN / "img_img_64_96.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    a = _img_args(tmp_path, meta)
    models.clear_model_cache()
    h0, m0 = models.cache_stats["hits"], models.cache_stats["misses"]
    pre = lambda k: image_ops.preprocess(torch.from_numpy(z[k]))
    outs = style.img_img_tensors(pre("content"), [pre("style1"), pre("style2")], a)
    # one checkpoint load + plan build for the whole schedule (the reference reloads per scale, optim.py:128-129)
    assert models.cache_stats["misses"] - m0 == 1 and models.cache_stats["hits"] - h0 == len(meta["sizes"]) - 1
    for size, out in zip(meta["sizes"], outs):
        assert out.is_cuda
        got = image_ops.deprocess_u8(out).cpu().numpy()
        ref = z[f"out_{size}"]
        assert got.shape == ref.shape
        p = psnr_u8(got, ref)
        print(f"img_img {size}px vs reference PNG: PSNR {p:.1f} dB")
        assert p > 38.0, (size, p)
    # the file-level entry writes the same PNGs
    from PIL import Image

    for k in ("content", "style1", "style2"):
        Image.fromarray(z[k], mode="RGB").save(tmp_path / f"{k}.png")
    a.content, a.style, a.output = str(tmp_path / "content.png"), [str(tmp_path / "style1.png"), str(tmp_path / "style2.png")], str(tmp_path / "out")
    style.img_img(a)
    for size, out in zip(meta["sizes"], outs):
        png = np.asarray(Image.open(tmp_path / f"out_{size}.png").convert("RGB"))
        assert np.array_equal(png, image_ops.deprocess_u8(out).cpu().numpy())
    assert style.img_img(a) == []  # resume: every scale already has its PNG (style.py:31-33)


def test_model_and_target_caches_do_not_change_results(tmp_path, monkeypatch):
    from maua_style_b200 import models, optim

    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    content = O.synthetic_image(64, 80, seed=1, smooth=True)
    styles = [O.synthetic_image(72, 64, seed=2), O.synthetic_image(56, 90, seed=3, smooth=True)]
    init = O.synthetic_image(64, 80, seed=4) * 0.25

    def run():
        a = make_args(ckpt, tmp_path, optimizer="adam", style_blend_weights=[0.75, 0.25])
        net, losses = models.load_model(a)
        outs = [optim.optimize(content, styles, init.clone(), 4, a, net, losses) for _ in range(2)]
        return outs, getattr(net, "style_cache_hits", 0)

    models.clear_model_cache()
    (a1, a2), hits = run()
    assert hits == 1                      # second optimize() call re-used the captured style targets
    assert torch.equal(a1, a2)            # ... and is bit-identical to the first (deterministic kernels)
    (b1, _), _ = run()                    # a new load_model on the cached core
    assert models.cache_stats["hits"] >= 1 and torch.equal(a1, b1)
    monkeypatch.setenv("MAUA_NO_MODEL_CACHE", "1")
    monkeypatch.setenv("MAUA_NO_TARGET_CACHE", "1")
    (c1, c2), hits = run()
    assert hits == 0 and torch.equal(a1, c1) and torch.equal(a1, c2)
    # an in-place edit of a style image invalidates the cached targets
    monkeypatch.delenv("MAUA_NO_TARGET_CACHE")
    a = make_args(ckpt, tmp_path, optimizer="adam", style_blend_weights=[0.75, 0.25])
    net, losses = models.load_model(a)
    optim.optimize(content, styles, init.clone(), 1, a, net, losses)
    t0 = net.style_losses[0].target.clone()
    styles[0].mul_(0.5)
    optim.optimize(content, styles, init.clone(), 1, a, net, losses)
    assert getattr(net, "style_cache_hits", 0) == 0 and not torch.equal(t0, net.style_losses[0].target)


def test_style_target_cache_follows_the_arithmetic_mode(tmp_path):
    """Targets captured by the TF32 kernels must not be re-used after net.set_impl() switched the plan to the exact-arithmetic
    kernels (and back): the cache signature carries the mode.  (__graft_entry__.smoke() switches modes on one net; with stale
    targets its exact-mode losses were 8e-4 away from the oracle.)"""
    from maua_style_b200 import _lib, models, optim

    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    style = [O.synthetic_image(64, 64, seed=2)]
    a = make_args(ckpt, tmp_path)
    net, losses = models.load_model(a)
    optim.set_style_targets(net, style, a)
    t_tf32 = [m.target.clone() for m in net.style_losses]
    optim.set_style_targets(net, style, a)
    assert getattr(net, "style_cache_hits", 0) == 1
    net.set_impl(_lib.MAUA_IMPL_FP32)
    optim.set_style_targets(net, style, a)
    assert net.style_cache_hits == 1                       # re-captured, not a cache hit
    t_exact = [m.target.clone() for m in net.style_losses]
    errs = [rel(x, y) for x, y in zip(t_tf32, t_exact)]
    assert any(e > 0 for e in errs) and max(errs) < 5e-3, errs
    net.set_impl(_lib.MAUA_IMPL_TC)
    optim.set_style_targets(net, style, a)
    assert net.style_cache_hits == 1
    for x, y in zip(t_tf32, (m.target for m in net.style_losses)):
        assert torch.equal(x, y)                           # deterministic kernels: the TF32 capture is reproduced bit for bit


def test_stylize_frame_matches_oracle(tmp_path):
    """One vid_img frame (style.py:276-296): warp -> temporal target with the resized reliability map -> blend -> Adam."""
    from maua_style_b200 import models, style

    ckpt = tmp_path / "vgg19-random.pth"
    params = save_checkpoint(ckpt)
    h, w = 72, 96
    content = O.synthetic_image(h, w, seed=1, smooth=True)
    styles = [O.synthetic_image(80, 88, seed=2)]
    prev = O.synthetic_image(h, w, seed=5, smooth=True)
    blend_img = O.synthetic_image(h, w, seed=6, smooth=True)
    flow = torch.from_numpy((np.random.RandomState(3).randn(36, 48, 2) * 0.02).astype(np.float32))
    reliable = torch.rand(1, 1, 36, 48, generator=torch.Generator().manual_seed(8))
    iters, tb = 5, 0.5
    a = make_args(ckpt, tmp_path, optimizer="adam", temporal_blend=tb)
    net, losses = models.load_model(a)
    from maua_style_b200 import image_ops

    grid = image_ops.flow_warp_grid(flow, (h, w))
    out = style.stylize_frame(net, losses, content, styles, a, iters, prev_pastiche=prev, flow_grid=grid,
                              reliable_flow=reliable, blend_image=blend_img).cpu()
    # oracle composition of the same steps
    fs = flow.numpy()
    neutral = np.rollaxis(np.array(np.meshgrid(np.linspace(-1, 1, 48), np.linspace(-1, 1, 36))), 0, 3)
    ogrid = I.resize_bilinear((neutral + fs).astype(np.float32).transpose(2, 0, 1)[None], size=(h, w))[0].transpose(1, 2, 0)
    assert np.array_equal(grid[0].cpu().numpy(), ogrid)
    warp = torch.from_numpy(I.grid_sample_border(prev[0].numpy(), ogrid))[None]
    rel_w = torch.from_numpy(I.resize_bilinear(reliable.numpy(), size=(h, w)))
    start = torch.from_numpy(I.blend(blend_img.numpy(), prev.numpy(), 1 - tb, tb))
    cfg = O.StyleConfig(content_weight=5.0, optimizer="adam")
    ref = O.optimize(content, styles, start, iters, cfg, params, temporal=(warp, rel_w))
    p = O.psnr(out, ref)
    print(f"stylize_frame vs oracle: PSNR {p:.1f} dB")
    assert p > 40.0, p
    # second frame on the same net: style targets are not re-captured
    style.stylize_frame(net, losses, content, styles, a, 1, prev_pastiche=out, flow_grid=grid, reliable_flow=reliable)
    assert net.style_cache_hits == 1


# ---- utils.match_histogram on the device ----
def test_match_histogram_vs_reference_outputs_and_oracle():
    """Against the outputs of the reference's own utils.match_histogram (tests/golden/make_golden_hist.py): the residual
    must be the reference's `1e-3 * randn` input noise pushed through the colour map, nothing more; against the
    float64 oracle (no noise): fp32 rounding of the affine map only."""
    from maua_style_b200 import image_ops
    from test_image_oracle import check_hist_match, hist_cases

    for name, target, sources, ref in hist_cases():
        y = image_ops.match_histogram(torch.from_numpy(target).cuda(), [torch.from_numpy(s).cuda() for s in sources], mode=True)
        y = y.cpu().numpy()
        check_hist_match(name, y, target, sources, ref)
        m, _ = I.match_histogram_affine(target, sources)
        tol = 4e-7 * float(np.abs(m).sum(axis=1).max()) * 260.0 + 1e-5  # |M| |x| eps_fp32 with a small constant
        want = I.match_histogram(target, sources)
        assert float(np.abs(y - want).max()) < tol, (name, float(np.abs(y - want).max()), tol)


def test_image_moments_exact_on_integers():
    """Integer-valued images: every sum and product is exact in fp64, so the moments must equal numpy's int64 sums."""
    from maua_style_b200 import image_ops

    for h, w, seed in [(37, 53, 1), (64, 64, 2), (1024, 1024, 3), (1, 1, 4), (3, 5, 5)]:
        rs = np.random.RandomState(seed)
        x = rs.randint(-124, 152, size=(1, 3, h, w)).astype(np.int64)
        m = image_ops.image_moments(torch.from_numpy(x.astype(np.float32)).cuda()).cpu().numpy()
        p = x.reshape(3, -1)
        want = [h * w, *p.sum(axis=1), (p[0] * p[0]).sum(), (p[0] * p[1]).sum(), (p[0] * p[2]).sum(), (p[1] * p[1]).sum(),
                (p[1] * p[2]).sum(), (p[2] * p[2]).sum()]
        assert np.array_equal(m, np.array(want, dtype=np.float64)), (h, w, m, want)


@pytest.mark.parametrize("h,w", [(1024, 1024), (2048, 2048), (1448, 1086), (1023, 769)])
def test_match_histogram_full_size_properties(h, w):
    """BASELINE.json's sizes: the matched image takes the sources' channel means and (one source) covariance; matching
    an image against itself is the identity; repeated launches are bit-identical (fixed-order reduction); the in-place
    form and pre-computed source moments give the same bits."""
    from maua_style_b200 import image_ops

    g = torch.Generator().manual_seed(h + w)
    low = torch.rand(1, 3, h // 8, w // 8, generator=g)
    t = (torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False) * 255 - 110).cuda()
    s = (torch.rand(1, 3, 700, 900, generator=g) * torch.tensor([200.0, 120.0, 60.0]).view(1, 3, 1, 1) - 90).cuda()
    y = image_ops.match_histogram(t, [s], mode=True)
    assert torch.equal(y, image_ops.match_histogram(t, [s], mode="avg"))
    sm = image_ops.image_moments(s).view(1, 10)
    y2 = image_ops.match_histogram(t.clone(), None, mode=True, source_moments=sm)
    assert torch.equal(y, y2)
    tc = t.clone()
    assert torch.equal(image_ops.match_histogram(tc, [s], mode=True, out=tc), y)
    my, ms = image_ops.image_moments(y).cpu().numpy(), sm[0].cpu().numpy()
    n = my[0]
    mean_y, mean_s = my[1:4] / n, ms[1:4] / ms[0]
    assert np.allclose(mean_y, mean_s, atol=2e-3), (mean_y, mean_s)
    var_y = my[[4, 7, 9]] / n - mean_y ** 2
    var_s = ms[[4, 7, 9]] / ms[0] - mean_s ** 2
    assert np.allclose(var_y, var_s, rtol=1e-3), (var_y, var_s)
    ident = image_ops.match_histogram(t, [t], mode=True)
    assert float((ident - t).abs().max()) < 2e-3
    assert image_ops.match_histogram(t, [s], mode=False) is t  # utils.py:97-98


def test_img_img_driver_with_histogram_matching(tmp_path):
    """style.py:24/:67/:71 with match_histograms on: every scale's result carries the style image's colour statistics."""
    from maua_style_b200 import image_ops, style

    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    args = make_args(ckpt, tmp_path, image_sizes=[64, 96], num_iters=[4, 3], init="content", style_scale=1.0, match_histograms=True)
    content = O.synthetic_image(80, 112, seed=1, smooth=True).cuda()
    sty = (O.synthetic_image(70, 90, seed=2, smooth=True) * 0.5 + 20).cuda()
    outs = style.img_img_tensors(content, [sty], args)
    ms = image_ops.image_moments(sty).cpu().numpy()
    for o in outs:
        mo = image_ops.image_moments(o).cpu().numpy()
        assert np.allclose(mo[1:4] / mo[0], ms[1:4] / ms[0], atol=1e-2)
        assert np.allclose(mo[[4, 7, 9]] / mo[0] - (mo[1:4] / mo[0]) ** 2, ms[[4, 7, 9]] / ms[0] - (ms[1:4] / ms[0]) ** 2, rtol=5e-3)


def test_loop_state_cache_reuses_buffers_and_graph_without_changing_results(tmp_path, monkeypatch):
    """optim.optimize_device keeps the optimizer state, the pastiche buffer and the captured CUDA graph per image size on the
    plan core (_LoopState): a batch of images of one size (shard.stylize_images, vid_img frames) must give exactly the
    results of fresh optimisations, L-BFGS and Adam, and the second image must replay the first image's graph."""
    from maua_style_b200 import models, optim

    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    style = [O.synthetic_image(64, 64, seed=2)]
    images = [O.synthetic_image(64, 80, seed=10 + i, smooth=True) for i in range(3)]
    for kind in ("lbfgs", "adam"):
        fresh = []
        monkeypatch.setenv("MAUA_NO_LOOP_CACHE", "1")
        for img in images:
            models.clear_model_cache()
            a = make_args(ckpt, tmp_path, optimizer=kind)
            net, losses = models.load_model(a)
            fresh.append(optim.optimize(img, style, img.clone(), 8, a, net, losses))
            del net, losses
        monkeypatch.setenv("MAUA_NO_LOOP_CACHE", "0")
        models.clear_model_cache()
        a = make_args(ckpt, tmp_path, optimizer=kind)
        net, losses = models.load_model(a)
        graphs = []
        for img, want in zip(images, fresh):
            got = optim.optimize(img, style, img.clone(), 8, a, net, losses)
            assert torch.equal(got, want), kind
            state = next(iter(net._core.loop_states.values()))
            graphs.append(state.iteration.graph)
        assert len(net._core.loop_states) == 1
        assert graphs[0] is not None and graphs[1] is graphs[0] and graphs[2] is graphs[0]  # captured once, replayed for all
        del net, losses


def test_host_pipelined_iteration_equals_sequential_steps(tmp_path):
    """optim.HostPipelinedIteration (bench.py's end-to-end path): every submit copies that step's host image to the device,
    runs one iteration on it and returns the PREVIOUS step's result from pinned host memory; with the copies on their own
    streams the results must be exactly those of doing the same steps one after the other."""
    from maua_style_b200 import models, optim

    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    a = make_args(ckpt, tmp_path, optimizer="adam")
    content = O.synthetic_image(64, 96, seed=1, smooth=True)
    style = [O.synthetic_image(64, 64, seed=2)]
    frames = [(O.synthetic_image(64, 96, seed=20 + i) * 0.25).pin_memory() for i in range(5)]

    def setup():
        net, losses = models.load_model(a)
        optim.set_content_targets(net, content, a)
        optim.set_style_targets(net, style, a)
        for m in losses:
            m.mode = "loss"
        p = frames[0].cuda().contiguous()
        opt = optim.PixelOptimizer(p, "adam", lr=1.0)
        up = torch.zeros(net._n_slots, device="cuda")
        up[net._live_slots()] = 1.0
        return net, losses, p, optim.GraphedIteration(net, p, opt, up)

    net, losses, p, it = setup()
    want = []
    for f in frames:
        p.copy_(f)
        it()
        want.append((p.detach().cpu().clone(), float(net._loss_vec.sum())))
    del net, losses, it
    net, losses, p, it = setup()
    pipe = optim.HostPipelinedIteration(it)
    got = []
    for f in frames:
        r = pipe.submit(f)
        if r is not None:
            got.append((r[0].clone(), float(r[1])))
    r = pipe.drain()
    got.append((r[0].clone(), float(r[1])))
    assert pipe.drain() is None and len(got) == len(frames)
    for (gi, gl), (wi, wl) in zip(got, want):
        assert torch.equal(gi, wi)
        assert abs(gl / wl - 1) < 1e-6
