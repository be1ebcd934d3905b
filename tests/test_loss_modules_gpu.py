"""The loss modules called on their own, outside a B200Net (the reference's `from loss import *` surface, loss.py:32-233: a user
can splice ContentLoss / StyleLoss / TVLoss into any nn.Sequential).  Values come from the library's kernels through the C ABI
(maua_content_loss_fwd, maua_tv_loss_fwd, maua_gram + maua_style_loss_fwd), gradients flow through torch autograd; checked
against the reference's formulas evaluated with torch CPU ops (fp32; fp64 for the Gram)."""
import pytest
import torch

from helpers import rel

pytestmark = pytest.mark.gpu


def seeded(shape, seed, scale=40.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("weighted", [False, True])
def test_standalone_content_loss(normalize, weighted):
    from maua_style_b200 import loss as L

    x0, tgt = seeded((1, 8, 20, 24), 1), seeded((1, 8, 20, 24), 2)
    w = torch.rand(1, 1, 20, 24, generator=torch.Generator().manual_seed(3)) if weighted else None
    strength = 5.0
    mod = L.ContentLoss(strength, normalize)
    mod.mode = "capture"
    mod(tgt.cuda())
    assert torch.equal(mod.target.cpu(), tgt) and mod.loss == 0
    mod.mode, mod.weights = "loss", (w.cuda() if weighted else None)
    x = x0.clone().cuda().requires_grad_(True)
    assert mod(x) is x  # identity on the features (loss.py:64)
    mod.loss.backward()
    # the reference's arithmetic with torch CPU ops
    xr = x0.clone().requires_grad_(True)
    mse = torch.nn.functional.mse_loss(xr * w, tgt) if weighted else torch.nn.functional.mse_loss(xr, tgt)
    (mse * strength).backward()
    want_loss, want_grad = float(mse.detach()) * strength, xr.grad * (strength if normalize else 1.0)  # ScaleGradients: strength^2 / |g| = strength
    assert abs(float(mod.loss) / want_loss - 1) < 1e-5
    assert rel(x.grad.cpu(), want_grad) < 1e-5
    # shape mismatch: silently skipped (loss.py:44); "none": untouched
    mod.loss = 0
    mod(torch.zeros(1, 8, 10, 12, device="cuda"))
    assert mod.loss == 0
    mod.mode = "none"
    mod(x)
    assert mod.loss == 0


def test_standalone_temporal_loss_without_target_is_skipped():
    from maua_style_b200 import loss as L

    mod = L.ContentLoss(50.0, True)
    mod.name = "temporal 1"
    mod.mode = "loss"
    mod(torch.zeros(1, 3, 16, 16, device="cuda"))
    assert mod.loss == 0  # loss.py:46-47


def test_standalone_tv_loss():
    from maua_style_b200 import loss as L

    x0 = seeded((1, 3, 37, 41), 5)
    x0[0, 0, 3, 4:8] = 7.0  # equal neighbours: sign(0) = 0
    mod = L.TVLoss(1e-3)
    x = x0.clone().cuda().requires_grad_(True)
    mod(x)
    mod.loss.backward()
    xr = x0.clone().requires_grad_(True)
    tv = 1e-3 * ((xr[:, :, 1:] - xr[:, :, :-1]).abs().sum() + (xr[:, :, :, 1:] - xr[:, :, :, :-1]).abs().sum())
    tv.backward()
    assert abs(float(mod.loss) / float(tv) - 1) < 1e-5
    assert rel(x.grad.cpu(), xr.grad) < 1e-6  # sums of +-1e-3
    assert float(x.grad[0, 0, 3, 5]) == float(xr.grad[0, 0, 3, 5])


@pytest.mark.parametrize("cov", [False, True])
def test_standalone_style_loss(cov):
    from maua_style_b200 import loss as L

    style, x0 = seeded((1, 64, 24, 28), 7), seeded((1, 64, 24, 28), 8)
    strength, vsf = 100.0, 100.0
    mod = L.StyleLoss(strength, use_covariance=cov, normalize=True, video_style_factor=vsf)
    mod.mode, mod.blend_weight = "capture", 1.0
    mod(style.cuda())
    mod.mode = "loss"
    x = x0.clone().cuda().requires_grad_(True)
    mod(x)
    mod.loss.backward()

    def gram(t):
        f = t.reshape(64, -1)
        if cov:
            f = f - f.mean(1, keepdim=True)
        return f @ f.T / t.nelement()

    xr = x0.clone().double().requires_grad_(True)
    mse = ((gram(xr) - gram(style.double())) ** 2).mean()
    mse.backward()
    # SURVEY 8a "net effect of R4-R6": value strength (1 + vsf) mse; gradient strength^2 (1 + [vsf > 0]) dmse/dF / |1| (normalised)
    assert abs(float(mod.loss) / (strength * (1 + vsf) * float(mse)) - 1) < 5e-3
    assert rel(x.grad.cpu(), (xr.grad * strength * strength * 2).float()) < 5e-3  # TF32 operands in the SYRK / the aux GEMM
