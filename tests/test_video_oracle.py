"""Pin the oracle's B > 1 (img_vid) semantics -- per-frame static Grams averaged over the window, the [B*C, B*C] dynamic
Gram, the per-term ScaleGradients, the window schedule and the overlap-frame gradient masking (loss.py:42-64, :141-181;
optim.py:69-90, :113-125, :149-170, :216-219) -- against goldens produced by the UNMODIFIED reference
(tests/golden/make_golden_arch.py: run_video_case).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import O, load_golden, video_cfg, video_inputs

CASES = ["img_vid_windows_adam_48x64", "img_vid_avgwin_lbfgs_48x48"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_window_feval_matches_reference(name):
    torch.set_flush_denormal(True)
    z, meta = load_golden(name)
    cfg = video_cfg(meta)
    content, styles, init = video_inputs(meta)
    gfw, afw = meta["gfw"], meta["afw"]
    net = O.OracleNet(O.he_init_vgg19(0), cfg)
    O.set_content_targets(net, content)
    first = styles if afw == -1 else [s[:afw] if s.shape[0] > 1 else s for s in styles]
    O.set_style_video_targets(net, first, cfg.blend(len(styles)), gfw)
    for m in net.losses:
        m.mode = "loss"
    for i, m in enumerate(net.style_losses):
        st = z[f"style_target_{i}_stats"]
        assert abs(m.target.norm().item() / st[1] - 1) < 1e-5
        np.testing.assert_allclose(m.target[:16, :16].numpy(), z[f"style_target_{i}_block"], rtol=1e-4, atol=1e-6 * st[1])
        assert list(m.video_target.shape) == list(z[f"video_target_{i}_shape"])
        assert abs(m.video_target.norm().item() / z[f"video_target_{i}_stats"][1] - 1) < 1e-5
    total, values, grad = O.feval(net, init[:gfw])
    keys = sorted([k for k in z.files if k.startswith("loss_")], key=lambda k: int(k.split("_")[1]))
    assert len(keys) == len(values)
    for k, v in zip(keys, values):
        assert abs(v - float(z[k])) <= 1e-5 * max(abs(float(z[k])), 1e-12), (k, v, float(z[k]))
    g_ref = torch.from_numpy(z["grad"])
    assert grad.shape == g_ref.shape and grad.shape[0] == gfw
    assert float((grad - g_ref).norm() / g_ref.norm()) < 1e-5


@pytest.mark.parametrize("name", CASES)
def test_oracle_windowed_optimisation_matches_reference(name):
    torch.set_flush_denormal(True)
    z, meta = load_golden(name)
    cfg = video_cfg(meta)
    content, styles, init = video_inputs(meta)
    out = O.optimize_windows(content, styles, init, meta["iters"], cfg, O.he_init_vgg19(0), meta["gfw"], meta["afw"])
    ref = torch.from_numpy(z["optimized"])
    assert out.shape == ref.shape
    assert O.psnr(out, ref) > 70.0, O.psnr(out, ref)
