"""Where the tolerances of the GPU parity tests come from (CPU only).

`oracle.tf32_emulation.TF32Net` is the fp32 oracle with the CUDA path's TF32 operand roundings inserted.  Its deviation
from the fp32 oracle on the golden inputs is what the sm_100a kernels are *expected* to show against the reference; the
numbers measured on the B200 (profiles/r01_parity_report.txt, profiles/r01f_parity_fullsize.txt) sit inside these
brackets, and the bounds asserted in tests/test_plan_gpu.py are the upper ends."""
import pytest
import torch

from helpers import O, golden_inputs, load_golden
from oracle.tf32_emulation import TF32Net, tf32_round


def nets(meta, pooling=None):
    over = dict(meta["over"])
    cfg = O.StyleConfig(content_weight=5.0)
    cfg.optimizer = over.pop("optimizer", "adam")
    cfg.normalize_gradients = not over.pop("no_grad_norm", False)
    if "style_blend_weights" in over:
        cfg.style_blend_weights = [float(x) for x in over.pop("style_blend_weights").split(",")]
    for k, v in over.items():
        setattr(cfg, k, v)
    if pooling:
        cfg.pooling = pooling
    params = O.he_init_vgg19(0)
    content, styles, init = golden_inputs(meta)
    out = []
    for cls in (O.OracleNet, TF32Net):
        net = cls(params, cfg)
        O.set_content_targets(net, content)
        O.set_style_targets(net, styles, cfg.blend(len(styles)))
        for m in net.losses:
            m.mode = "loss"
        out.append(net)
    return out, (content, styles, init), cfg, params


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_tf32_rounding_is_cvt_rna():
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -11 + 2 ** -20, -1.0 - 2 ** -11, 3.14159265, 0.0, 1e-30])
    r = tf32_round(x)
    assert r[0] == 1.0 and r[1] == 1.0 + 2 ** -10 and r[2] == 1.0 + 2 ** -10 and r[3] == -1.0 - 2 ** -10  # ties away
    assert abs(float(r[4]) / 3.14159265 - 1) <= 2 ** -11 and r[5] == 0.0
    assert torch.equal(tf32_round(r), r)


@pytest.mark.parametrize("name,pooling", [("adam_gram_90x122", "max"), ("adam_gram_90x122", "avg"),
                                          ("adam_cov_2styles_96x128", "max")])
def test_expected_feature_loss_and_gradient_deviation(name, pooling):
    torch.set_flush_denormal(True)
    _, meta = load_golden(name)
    (ref, emu), (content, styles, init), cfg, _ = nets(meta, pooling)
    taps_r, taps_e = {}, {}
    with torch.no_grad():
        ref(init.clone(), taps=taps_r)
        emu(init.clone(), taps=taps_e)
    for m in ref.losses + emu.losses:
        m.loss = 0
    # features: eps_tf32 = 2^-11 per rounding, growing ~ sqrt(depth); GPU tests bound them at 1e-3 / 2e-3
    for nm in ("relu1_1", "relu3_1", "relu5_1"):
        err = rel(taps_e[nm], taps_r[nm])
        assert 5e-5 < err < (1e-3 if nm != "relu5_1" else 2e-3), (nm, err)
    for a, b in zip(emu.style_losses, ref.style_losses):
        assert rel(a.target, b.target) < 2e-3
    _, vr, gr = O.feval(ref, init)
    _, ve, ge = O.feval(emu, init)
    for a, b in zip(ve, vr):
        if b != 0.0:
            assert abs(a / b - 1) < 1e-2
    gerr = rel(ge, gr)
    if pooling == "avg":
        assert gerr < 3e-3, gerr           # arithmetic error only
    else:
        assert 3e-3 < gerr < 4e-2, gerr    # arg-max / ReLU-sign flips dominate: an order of magnitude above the avg case


def test_expected_image_deviation_after_short_optimisations():
    """PSNR bound of the optimize tests (>= 40 dB after <= 10 iterations): the emulated path lands where the B200 did
    (58.1 dB measured for lbfgs_gram_64, 49.7 dB for adam_gram_64)."""
    torch.set_flush_denormal(True)
    for name, lo in (("lbfgs_gram_64", 50.0), ("adam_gram_64", 44.0)):
        z, meta = load_golden(name)
        (_, emu), (content, styles, init), cfg, params = nets(meta)

        def closure(p):
            return O.feval(emu, p)[2]

        if cfg.optimizer == "adam":
            out = O.adam_optimize(init.clone(), closure, meta["iters"] + 1, lr=cfg.learning_rate)
        else:
            out = O.lbfgs_optimize(init.clone(), closure, meta["iters"], lr=1.0, history=cfg.lbfgs_num_correction)
        p = O.psnr(out, torch.from_numpy(z["optimized"]))
        assert p > lo, (name, p)


def test_adam_loss_trajectory_is_insensitive_to_tf32_rounding():
    """30 Adam iterations: the loss curve of the emulated CUDA arithmetic follows the fp32 one to < 1 % although individual
    pixels drift apart (every Adam step moves every pixel by ~lr) -- why long runs are compared by loss, not by PSNR."""
    torch.set_flush_denormal(True)
    _, meta = load_golden("adam_gram_64")
    (ref, emu), (content, styles, init), cfg, _ = nets(meta)
    curves = []
    for net in (ref, emu):
        hist = []

        def closure(p, net=net, hist=hist):
            tot, _, g = O.feval(net, p)
            hist.append(tot)
            return g

        O.adam_optimize(init.clone(), closure, 30, lr=1.0)
        curves.append(hist)
    for a, b in zip(*curves):
        assert abs(b / a - 1) < 1e-2
    assert curves[0][-1] < 0.7 * curves[0][0]


def test_expected_deviation_of_the_video_driver():
    """The vid_img driver chains 12 short optimisations (2 scales x 2 passes x 3 frames), every frame starting from earlier
    results, so TF32 rounding accumulates along the chain.  With the CUDA path's operand roundings inserted on the CPU, the oracle's
    driver lands where the B200 measured the device driver (46.3 dB worst frame, profiles/r05_vid_driver.txt): the bound of
    tests/test_vid_driver_gpu.py (40 dB) is rounding, not a driver difference -- in exact arithmetic the same driver gives 52.7 dB."""
    import json

    import numpy as np

    from helpers import GOLDEN
    from oracle import image_oracle as I

    z = np.load(GOLDEN / "vid_img_3f_48_80.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    params = O.he_init_vgg19(0)
    cfg = O.StyleConfig(content_weight=meta["content_weight"], style_weight=meta["style_weight"], tv_weight=meta["tv_weight"],
                        temporal_weight=meta["temporal_weight"], optimizer=meta["optimizer"])
    torch.set_flush_denormal(True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))

    def optimize_fn(content, styles, pastiche, iters, temporal):  # O.optimize with the emulated network
        net = TF32Net(params, cfg)
        O.set_content_targets(net, t(content))
        if temporal is not None:
            O.set_temporal_targets(net, t(temporal[0]), t(temporal[1]))
        O.set_style_targets(net, [t(s) for s in styles], cfg.blend(len(styles)))
        for m in net.losses:
            m.mode = "loss"
        return O.adam_optimize(t(pastiche).clone(), lambda p: O.feval(net, p)[2], iters + 1, lr=cfg.learning_rate).detach().numpy()

    frames = [z[f"frame_{i}"] for i in range(meta["n_frames"])]
    store = I.vid_img(frames, [I.preprocess_u8(z["style"])], meta["sizes"], meta["iters"], meta["passes"], optimize_fn,
                      lambda d, a, b: (z[f"flow_{d}_{a}_{b}"], z[f"rel_{d}_{a}_{b}"]), init=meta["init"],
                      temporal_blend=meta["temporal_blend"])
    worst = 99.0
    for (size, p, f), got in store.items():
        ref = z[f"out_{size}_{p}_{f}"]
        mse = float(((got.astype(np.float64) - ref.astype(np.float64)) ** 2).mean())
        worst = min(worst, 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse))
    print(f"vid_img driver under TF32 emulation: worst frame {worst:.1f} dB (B200: 46.3 dB)")
    assert 43.0 < worst < 50.0, worst  # emulated: 46.1 dB


def test_expected_deviation_of_the_frame_window_driver(monkeypatch):
    """The img_vid driver (2 scales, 4 + 6 windows of 3 / 2 frames) with the CUDA path's operand roundings inserted on the CPU:
    56.8 / 51.0 dB at 32 / 48 px -- the B200 measured 56.6 / 51.0 dB (profiles/r05_vid_driver.txt); exact arithmetic: 168.6 / 94.4 dB."""
    import json

    import numpy as np

    from helpers import GOLDEN
    from oracle import image_oracle as I

    z = np.load(GOLDEN / "img_vid_9f_32_48.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    params = O.he_init_vgg19(0)
    cfg = O.StyleConfig(content_weight=meta["content_weight"], style_weight=meta["style_weight"], tv_weight=meta["tv_weight"],
                        video_style_factor=meta["video_style_factor"], optimizer=meta["optimizer"])
    torch.set_flush_denormal(True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    monkeypatch.setattr(O, "OracleNet", TF32Net)  # optimize_windows builds its network through this name
    outs = I.img_vid(I.preprocess_u8(z["content"]), [z["style_clip"]], z["init_video"], meta["sizes"], meta["iters"],
                     [int(w) for w in meta["windows"].split(",")],
                     lambda c, s, p, it, gfw: O.optimize_windows(t(c), [t(x) for x in s], t(p), it, cfg, params, gfw=gfw).detach().numpy(),
                     temporal_blend=meta["temporal_blend"])
    for (size, lo, hi), out in zip(((32, 52.0, 62.0), (48, 47.0, 56.0)), outs):
        p = O.psnr(t(out), t(z[f"out_{size}"]))
        print(f"img_vid driver under TF32 emulation, {size}px: {p:.1f} dB")
        assert lo < p < hi, (size, p)
