"""Per-kernel parity of the CUDA path (through the C ABI) against plain fp32 torch on the CPU.

Floating-point kernels: the conv / Gram kernels compute with TF32 operands and FP32 accumulation, so the
tolerance is stated per test as a relative L2 error (||a-b|| / ||b||); element-wise kernels are compared at
fp32 round-off.  The naive SIMT implementation (MAUA_IMPL_REF) is cross-checked too: it shares the
contract, not the code, of the tcgen05 path.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from maua_style_b200 import _lib

pytestmark = pytest.mark.gpu

TF32_REL = 1.5e-3  # single conv layer, TF32 operands (eps 2^-11 each), fp32 accumulate


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def tf32_round(x):
    """RNA rounding of fp32 to tf32 (10-bit mantissa), matching cvt.rna.tf32.f32."""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


IMPLS = pytest.mark.parametrize("impl", [pytest.param(_lib.MAUA_IMPL_REF, id="ref"), pytest.param(_lib.MAUA_IMPL_TC, id="tc"),
                                         pytest.param(_lib.MAUA_IMPL_TC_1CTA, id="tc1cta"),
                                         pytest.param(_lib.MAUA_IMPL_TC_2CTA, id="tc2cta")])


@pytest.fixture(scope="module")
def lib():
    _lib.require_gpu()
    return _lib.load()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def prep(lib, w, dgrad):
    cout, cin = w.shape[:2]
    out = torch.empty((cin if dgrad else cout, 9 * (cout if dgrad else cin)), device="cuda")
    _lib.check(lib.maua_prep_conv_weights(_lib.ptr(w), _lib.ptr(out), cout, cin, int(dgrad), _lib.stream_ptr()))
    return out


@IMPLS
@pytest.mark.parametrize(
    "cin,cout,h,w",
    [(64, 64, 32, 48), (64, 128, 16, 16), (128, 256, 24, 40), (256, 512, 11, 13), (512, 512, 8, 8), (64, 64, 90, 122),
     (128, 128, 181, 77)],
)
def test_conv3x3_fwd(lib, impl, cin, cout, h, w):
    g = torch.Generator().manual_seed(cin * 7 + cout + h)
    x = tf32_round(torch.randn(1, cin, h, w, generator=g))
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(x, tf32_round(wt), b, padding=1))
    xd, wd, bd = nhwc(x).cuda(), wt.cuda(), b.cuda()
    wg = prep(lib, wd, False)
    y = torch.empty(1, h, w, cout, device="cuda")
    _lib.check(lib.maua_conv3x3_fwd(_lib.ptr(xd), _lib.ptr(wg), _lib.ptr(bd), _lib.ptr(y), 1, h, w, cin, cout, 1, impl,
                                    _lib.stream_ptr()), "conv3x3_fwd")
    torch.cuda.synchronize()
    err = rel(nchw(y), ref)
    exact = tf32_round(F.relu(F.conv2d(x.double(), tf32_round(wt).double(), b.double(), padding=1)).float())
    print(f"conv fwd impl={impl} {cin}->{cout} {h}x{w}: rel vs fp32 torch {err:.2e}, vs fp64-exact-then-rounded {rel(nchw(y), exact):.2e}")
    assert err < TF32_REL, f"conv fwd rel err {err}"


@IMPLS
@pytest.mark.parametrize("cin,cout,h,w", [(64, 64, 32, 48), (64, 128, 17, 23), (256, 512, 12, 20), (512, 512, 9, 7)])
def test_conv3x3_dgrad_with_mask(lib, impl, cin, cout, h, w):
    g = torch.Generator().manual_seed(cin + cout * 3 + h)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    gy = tf32_round(torch.randn(1, cout, h, w, generator=g))
    act = torch.randn(1, cin, h, w, generator=g)  # sign pattern of the layer below
    x = torch.zeros(1, cin, h, w, requires_grad=True)
    F.conv2d(x, tf32_round(wt), None, padding=1).backward(gy)
    ref = x.grad * (act > 0)
    wtd = wt.cuda()
    wdg = prep(lib, wtd, True)
    gx = torch.empty(1, h, w, cin, device="cuda")
    null = C.c_void_p(0)
    gyd, actd = nhwc(gy).cuda(), nhwc(act).cuda()  # keep device tensors alive until the sync
    _lib.check(lib.maua_conv3x3_dgrad(_lib.ptr(gyd), _lib.ptr(wdg), _lib.ptr(gx), 1, h, w, cout, cin,
                                      _lib.ptr(actd), null, null, null, null, null, null, 0, impl,
                                      _lib.stream_ptr()), "conv3x3_dgrad")
    torch.cuda.synchronize()
    err = rel(nchw(gx), ref)
    assert err < TF32_REL, f"dgrad rel err {err}"


@IMPLS
@pytest.mark.parametrize("cin,cout,h,w", [(64, 64, 32, 48), (64, 128, 17, 23), (256, 512, 12, 20), (512, 256, 9, 7)])
def test_conv3x3_dgrad_with_bitmap_mask(lib, impl, cin, cout, h, w):
    """Same contract as above with the ReLU mask as a sign bitmap (what the plan uses): the tcgen05 path then takes the
    TMA-store epilogue.  Also checks maua_relu_mask_bits itself, bit for bit."""
    g = torch.Generator().manual_seed(cin * 5 + cout + w)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    gy = tf32_round(torch.randn(1, cout, h, w, generator=g))
    act = torch.randn(1, cin, h, w, generator=g)
    x = torch.zeros(1, cin, h, w, requires_grad=True)
    F.conv2d(x, tf32_round(wt), None, padding=1).backward(gy)
    ref = x.grad * (act > 0)
    wdg = prep(lib, wt.cuda(), True)
    gyd, actd = nhwc(gy).cuda(), nhwc(act).cuda()
    bits = torch.zeros(h * w * cin // 32, dtype=torch.int32, device="cuda")
    _lib.check(lib.maua_relu_mask_bits(_lib.ptr(actd), _lib.ptr(bits), C.c_long(h * w), cin, _lib.stream_ptr()), "mask_bits")
    torch.cuda.synchronize()
    want = (nhwc(act).reshape(-1, 32) > 0).to(torch.int64)
    want = (want << torch.arange(32)).sum(1)
    got = bits.cpu().to(torch.int64) & 0xFFFFFFFF
    assert torch.equal(got, want)
    gx = torch.empty(1, h, w, cin, device="cuda")
    null = C.c_void_p(0)
    _lib.check(lib.maua_conv3x3_dgrad_bits(_lib.ptr(gyd), _lib.ptr(wdg), _lib.ptr(gx), 1, h, w, cout, cin, _lib.ptr(bits),
                                           null, null, null, 0, impl, _lib.stream_ptr()), "conv3x3_dgrad_bits")
    torch.cuda.synchronize()
    err = rel(nchw(gx), ref)
    assert err < TF32_REL, f"dgrad (bitmap mask) rel err {err}"


@IMPLS
@pytest.mark.parametrize("with_main", [True, False])
def test_dgrad_with_style_and_content_terms(lib, impl, with_main):
    """gx = (dgrad(gy) + F @ D + bias + coef (F - T)) * (F > 0)  -- the fused tap-gradient epilogue."""
    cin, cout, h, w = 128, 256, 20, 28
    g = torch.Generator().manual_seed(5)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.03
    gy = tf32_round(torch.randn(1, cout, h, w, generator=g))
    feat = tf32_round(torch.randn(1, cin, h, w, generator=g))
    targ = torch.randn(1, cin, h, w, generator=g)
    D = torch.randn(cin, cin, generator=g) * 0.05
    D = tf32_round(D + D.t())
    bias = torch.randn(cin, generator=g)
    coef = torch.tensor([0.37])
    ref = torch.einsum("bdhw,cd->bchw", feat, D) + bias.view(1, -1, 1, 1) + coef * (feat - targ)
    if with_main:
        x = torch.zeros(1, cin, h, w, requires_grad=True)
        F.conv2d(x, tf32_round(wt), None, padding=1).backward(gy)
        ref = ref + x.grad
    ref = ref * (feat > 0)
    gx = torch.empty(1, h, w, cin, device="cuda")
    fd, td = nhwc(feat).cuda(), nhwc(targ).cuda()
    null = C.c_void_p(0)
    gyd = nhwc(gy).cuda() if with_main else None
    wtd = wt.cuda()
    wdg = prep(lib, wtd, True) if with_main else None
    Dd, biasd, coefd = D.cuda(), bias.cuda(), coef.cuda()
    _lib.check(lib.maua_conv3x3_dgrad(_lib.ptr(gyd), _lib.ptr(wdg), _lib.ptr(gx), 1, h, w, cout, cin, _lib.ptr(fd),
                                      _lib.ptr(fd), _lib.ptr(Dd), _lib.ptr(biasd), _lib.ptr(fd), _lib.ptr(td),
                                      _lib.ptr(coefd), 0, impl, _lib.stream_ptr()), "conv3x3_dgrad+aux")
    torch.cuda.synchronize()
    err = rel(nchw(gx), ref)
    assert err < TF32_REL, f"fused dgrad rel err {err}"


@pytest.mark.parametrize("h,w", [(32, 32), (37, 53), (5, 3)])
def test_conv_first_fwd_and_dgrad(lib, h, w):
    g = torch.Generator().manual_seed(h)
    img = torch.rand(1, 3, h, w, generator=g) * 255 - 110
    wt = torch.randn(64, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(64, generator=g)
    ref = F.relu(F.conv2d(img, wt, b, padding=1))
    y = torch.empty(1, h, w, 64, device="cuda")
    imgd, wtd, bd = img.cuda(), wt.cuda(), b.cuda()
    _lib.check(lib.maua_conv_first_fwd(_lib.ptr(imgd), _lib.ptr(wtd), _lib.ptr(bd), _lib.ptr(y), 1, h, w,
                                       64, _lib.stream_ptr()))
    torch.cuda.synchronize()
    # tcgen05 kernel with 3xTF32 operand splitting: ~1e-6 before the output is rounded to TF32, so a fraction of a percent of
    # the outputs lands on the other side of a TF32 rounding boundary (one such flip = 2^-11 of that element)
    assert rel(nchw(y), tf32_round(ref)) < 1e-4

    # dgrad + TV + temporal tail
    gy = tf32_round(torch.randn(1, 64, h, w, generator=g))  # the plan hands over a TF32-rounded masked gradient
    x = img.clone().requires_grad_(True)
    warp = torch.rand(1, 3, h, w, generator=g) * 255 - 110
    wts = torch.rand(1, 1, h, w, generator=g)
    tvc, tpc = 0.013, 0.7
    tv = (x[:, :, 1:, :] - x[:, :, :-1, :]).abs().sum() + (x[:, :, :, 1:] - x[:, :, :, :-1]).abs().sum()
    temporal = ((x * wts - warp) ** 2).sum() * 0.5
    total = (F.conv2d(x, wt, None, padding=1) * gy).sum() + tvc * tv + tpc * temporal
    total.backward()
    gimg = torch.empty(1, 3, h, w, device="cuda")
    gyd, warpd, wtsd = nhwc(gy).cuda(), warp.cuda(), wts.cuda()
    tvd, tpd = torch.tensor([tvc]).cuda(), torch.tensor([tpc]).cuda()
    ws = torch.empty(lib.maua_conv_first_dgrad_workspace_bytes(1, h, w), dtype=torch.uint8, device="cuda")
    _lib.check(lib.maua_conv_first_dgrad(_lib.ptr(gyd), _lib.ptr(wtd), _lib.ptr(gimg), 1, h, w, 64,
                                         _lib.ptr(imgd), _lib.ptr(tvd), _lib.ptr(warpd),
                                         _lib.ptr(wtsd), _lib.ptr(tpd), _lib.ptr(ws), _lib.stream_ptr()))
    torch.cuda.synchronize()
    # the 64 -> 27 channel contraction runs on the tensor core with TF32 operands (gy, w rounded to 10-bit mantissas)
    assert rel(gimg, x.grad) < 1e-3


@pytest.mark.parametrize("avg", [0, 1])
@pytest.mark.parametrize("h,w", [(16, 16), (17, 23), (9, 2)])
def test_pool_fwd_bwd(lib, avg, h, w):
    c = 64
    g = torch.Generator().manual_seed(h * w)
    x = F.relu(torch.randn(1, c, h, w, generator=g))
    x[0, :, : h // 2 * 2 : 2, : w // 2 * 2 : 2] = x[0, :, 1 : h // 2 * 2 : 2, 1 : w // 2 * 2 : 2]  # force ties
    x = tf32_round(x).requires_grad_(True)
    y = (F.avg_pool2d if avg else F.max_pool2d)(x, 2, 2)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    ref_g = x.grad * (x > 0)
    yd = torch.empty(1, h // 2, w // 2, c, device="cuda")
    xd = nhwc(x.detach()).cuda()
    _lib.check(lib.maua_pool2x2_fwd(_lib.ptr(xd), _lib.ptr(yd), 1, h, w, c, avg, _lib.stream_ptr()))
    gx = torch.full((1, h, w, c), 7.0, device="cuda")
    gyd = nhwc(gy).cuda()
    _lib.check(lib.maua_pool2x2_bwd(_lib.ptr(xd), _lib.ptr(gyd), C.c_void_p(0), _lib.ptr(gx), 1, h, w, c, avg, 0,
                                    _lib.stream_ptr()))
    torch.cuda.synchronize()
    yref = y.detach()
    assert rel(nchw(yd), tf32_round(yref) if avg else yref) < 2e-5
    assert rel(nchw(gx), ref_g) < 1e-6


@IMPLS
@pytest.mark.parametrize("cov", [0, 1])
@pytest.mark.parametrize("c,h,w", [(64, 40, 56), (128, 33, 21), (256, 16, 16), (512, 9, 13), (64, 3, 5),
                                   (192, 24, 20), (320, 12, 17), (1536, 8, 9)])  # B*C of img_vid windows; pruned VGG-16 (192, 320)
def test_gram(lib, impl, cov, c, h, w):
    g = torch.Generator().manual_seed(c + h)
    f = tf32_round(F.relu(torch.randn(1, c, h, w, generator=g) + 0.3))
    X = f.reshape(c, h * w).double()
    if cov:
        X = X - X.mean(1, keepdim=True)
    ref = (X @ X.t() / (c * h * w)).float()
    fd = nhwc(f).cuda()
    ws = torch.empty(lib.maua_gram_workspace_bytes(c), dtype=torch.uint8, device="cuda")
    G = torch.empty(c, c, device="cuda")
    mean = torch.empty(c, device="cuda")
    _lib.check(lib.maua_gram(_lib.ptr(fd), C.c_long(h * w), c, cov, _lib.ptr(G), _lib.ptr(mean), _lib.ptr(ws), impl,
                             _lib.stream_ptr()), "gram")
    torch.cuda.synchronize()
    err = rel(G, ref)
    assert err < (2e-5 if cov else 2e-6), f"gram rel err {err}"
    assert rel(G, G.t()) < 1e-6


def test_style_loss_and_prep(lib):
    c, p = 128, 1000
    g = torch.Generator().manual_seed(0)
    G, A = torch.randn(c, c, generator=g), torch.randn(c, c, generator=g)
    mean = torch.rand(c, generator=g)
    ws = torch.zeros(lib.maua_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
    loss = torch.zeros(1, device="cuda")
    diff = torch.empty(c, c, device="cuda")
    Gd, Ad, meand = G.cuda(), A.cuda(), mean.cuda()
    _lib.check(lib.maua_style_loss_fwd(_lib.ptr(Gd), _lib.ptr(Ad), c, C.c_float(3.5), _lib.ptr(loss),
                                       _lib.ptr(diff), _lib.ptr(ws), _lib.stream_ptr()))
    coef = torch.tensor([2.25], device="cuda")
    D = torch.empty(c, c, device="cuda")
    bias = torch.empty(c, device="cuda")
    _lib.check(lib.maua_style_loss_bwd_prep(_lib.ptr(diff), _lib.ptr(meand), c, C.c_long(p), _lib.ptr(coef),
                                            _lib.ptr(D), _lib.ptr(bias), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert abs(loss.item() - 3.5 * ((G - A) ** 2).mean().item()) < 1e-4
    Dref = tf32_round(2.25 * 4 / (c ** 3 * p) * (G - A))
    assert rel(D, Dref) < 1e-6
    assert rel(bias, -(Dref @ mean)) < 1e-5


def test_content_tv_adam(lib):
    g = torch.Generator().manual_seed(1)
    ws = torch.zeros(lib.maua_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
    x, t = torch.randn(4096 * 3, generator=g), torch.randn(4096 * 3, generator=g)
    loss = torch.zeros(1, device="cuda")
    xd, td = x.cuda(), t.cuda()
    _lib.check(lib.maua_content_loss_fwd(_lib.ptr(xd), C.c_void_p(0), _lib.ptr(td), C.c_long(x.numel()),
                                         C.c_long(0), C.c_float(5.0), _lib.ptr(loss), _lib.ptr(ws), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert abs(loss.item() / (5.0 * F.mse_loss(x, t).item()) - 1) < 1e-5
    wts = torch.rand(4096, generator=g)
    wtsd = wts.cuda()
    _lib.check(lib.maua_content_loss_fwd(_lib.ptr(xd), _lib.ptr(wtsd), _lib.ptr(td), C.c_long(x.numel()),
                                         C.c_long(4096), C.c_float(50.0), _lib.ptr(loss), _lib.ptr(ws), _lib.stream_ptr()))
    torch.cuda.synchronize()
    refw = 50.0 * F.mse_loss(x.view(3, 4096) * wts, t.view(3, 4096)).item()
    assert abs(loss.item() / refw - 1) < 1e-5

    img = torch.rand(1, 3, 37, 41, generator=g) * 255
    imgd = img.cuda()
    _lib.check(lib.maua_tv_loss_fwd(_lib.ptr(imgd), 3, 37, 41, C.c_float(1e-3), _lib.ptr(loss), _lib.ptr(ws),
                                    _lib.stream_ptr()))
    torch.cuda.synchronize()
    tv = 1e-3 * ((img[:, :, 1:] - img[:, :, :-1]).abs().sum() + (img[:, :, :, 1:] - img[:, :, :, :-1]).abs().sum())
    assert abs(loss.item() / tv.item() - 1) < 1e-5

    n = 3 * 37 * 41 + 3
    p = torch.randn(n, generator=g).requires_grad_(True)
    opt = torch.optim.Adam([p], lr=1.0)
    pd = p.detach().clone().cuda()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 6):
        grad = torch.randn(n, generator=g) * 10 ** (step - 3)
        p.grad = grad.clone()
        opt.step()
        gradd = grad.cuda()
        _lib.check(lib.maua_adam_step(_lib.ptr(pd), _lib.ptr(gradd), _lib.ptr(m), _lib.ptr(v), C.c_long(n),
                                      C.c_float(1.0), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8), step,
                                      _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert rel(pd, p.detach()) < 1e-6


@pytest.mark.parametrize("n,history,iters", [(4096, 5, 30), (1003 * 4, 100, 40), (1003 * 4, 70, 85), (3 * 37 * 5, 7, 25),
                                             (3 * 101 * 67 + 1, 12, 16)])
def test_lbfgs_matches_torch(lib, n, history, iters):
    """Device-resident L-BFGS vs torch.optim.LBFGS driven the way optim.py:180-191 drives it."""
    g = torch.Generator().manual_seed(n)
    scale = torch.rand(n, generator=g) * 0.6 + 0.7
    shift = torch.randn(n, generator=g)

    def f(x):  # smooth, convex, non-quadratic, well conditioned (L-BFGS without line search converges)
        z = (x - shift) * scale
        return (0.5 * z * z + 0.02 * z ** 4).sum() / n

    x0 = torch.randn(n, generator=g)
    p = x0.clone().requires_grad_(True)
    opt = torch.optim.LBFGS([p], max_iter=iters, history_size=history, tolerance_change=-1, tolerance_grad=-1)

    def closure():
        opt.zero_grad()
        l = f(p)
        l.backward()
        return l

    opt.step(closure)

    state = C.c_void_p()
    _lib.check(lib.maua_lbfgs_create(C.c_long(n), history, C.c_float(1.0), C.c_float(-1.0), C.byref(state)))
    xd = x0.clone().cuda()
    for _ in range(iters):
        xh = xd.cpu().requires_grad_(True)
        f(xh).backward()
        gd = xh.grad.cuda()
        _lib.check(lib.maua_lbfgs_step(state, _lib.ptr(xd), _lib.ptr(gd), _lib.stream_ptr()))
        torch.cuda.synchronize()
    n_iter, hist, halted = C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.maua_lbfgs_query(state, C.byref(n_iter), C.byref(hist), C.byref(halted), _lib.stream_ptr()))
    lib.maua_lbfgs_destroy(state)
    # the y.s > 1e-10 gate rejects pairs once the iteration has converged: the history length must match torch's
    torch_hist = len(opt.state_dict()["state"][0]["old_dirs"])
    assert n_iter.value == iters and halted.value == 0 and hist.value == torch_hist, (n_iter.value, hist.value, torch_hist)
    assert f(xd.cpu()).item() <= f(x0).item()
    assert rel(xd, p.detach()) < 1e-4, f"lbfgs rel err {rel(xd, p.detach())}"


# ---- NIN layer shapes (models.py:74-113): direct convolutions and the 3x3 / 2 ceil_mode pools (csrc/conv_gen.cu) ----------------
@pytest.mark.parametrize("cin,cout,ks,stride,pad,h,w,img", [(3, 96, 11, 4, 0, 67, 90, True), (3, 128, 11, 4, 0, 131, 150, True),
                                                            (96, 256, 5, 1, 2, 15, 18, False), (128, 256, 5, 1, 2, 33, 9, False)])
def test_direct_conv_forward_and_input_gradient(lib, cin, cout, ks, stride, pad, h, w, img):
    g = torch.Generator().manual_seed(cin + h)
    x = torch.randn(1, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, ks, ks, generator=g) * (2.0 / (cin * ks * ks)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    xr = x.double().requires_grad_(True)
    ref = F.relu(F.conv2d(xr, wt.double(), b.double(), stride=stride, padding=pad))
    oh, ow = ref.shape[2], ref.shape[3]
    out = torch.empty(1, oh, ow, cout, device="cuda")
    xin = x.cuda().contiguous() if img else nhwc(x).cuda()
    wd, bd = wt.cuda(), b.cuda()  # (named: a temporary's block would be handed to the next temporary while the kernel is pending)
    _lib.check(lib.maua_conv_direct_fwd(_lib.ptr(xin), int(img), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(out), 1, h, w, cin,
                                        cout, ks, stride, pad, 1, 0, _lib.stream_ptr()), "conv_direct_fwd")
    torch.cuda.synchronize()
    assert rel(nchw(out), ref.float()) < 2e-6
    # input gradient for a random output gradient
    go = torch.randn(1, cout, oh, ow, generator=g)
    pre = F.conv2d(xr, wt.double(), b.double(), stride=stride, padding=pad)
    (gref,) = torch.autograd.grad(pre, xr, go.double())
    god = nhwc(go).cuda()
    if img:
        gx = torch.empty(1, cin, h, w, device="cuda")
        _lib.check(lib.maua_conv_direct_dgrad_image(_lib.ptr(god), _lib.ptr(wd), _lib.ptr(gx), 1, h, w, cout, ks, stride,
                                                    _lib.stream_ptr()), "conv_direct_dgrad_image")
        torch.cuda.synchronize()
        assert rel(gx, gref.float()) < 2e-6
    else:
        wf = torch.empty(cin, cout, ks, ks, device="cuda")
        _lib.check(lib.maua_conv_direct_flip_weights(_lib.ptr(wd), _lib.ptr(wf), cout, cin, ks, _lib.stream_ptr()))
        gx = torch.empty(1, h, w, cin, device="cuda")
        _lib.check(lib.maua_conv_direct_fwd(_lib.ptr(god), 0, _lib.ptr(wf), C.c_void_p(0), _lib.ptr(gx), 1, oh, ow, cout, cin, ks, 1,
                                            ks - 1 - pad, 0, 0, _lib.stream_ptr()), "conv_direct dgrad")
        torch.cuda.synchronize()
        assert rel(nchw(gx), gref.float()) < 2e-6


@pytest.mark.parametrize("avg", [0, 1])
@pytest.mark.parametrize("c,h,w", [(96, 30, 37), (128, 15, 15), (256, 7, 8), (64, 2, 5), (32, 3, 4)])
def test_pool3x3_ceil_forward_backward(lib, avg, c, h, w):
    """MaxPool2d / AvgPool2d((3,3),(2,2),(0,0), ceil_mode=True) and their backward with the ReLU mask of the producer folded in
    (models.py:77-80); ties: quantised inputs make equal maxima frequent, torch's CPU kernel is the arbiter."""
    g = torch.Generator().manual_seed(c * 7 + h)
    x = F.relu(torch.round(torch.randn(1, c, h, w, generator=g) * 2) / 2)
    xr = x.clone().requires_grad_(True)
    pool = (F.avg_pool2d if avg else F.max_pool2d)
    y = pool(F.relu(xr), 3, 2, 0, ceil_mode=True)
    ph, pw = y.shape[2], y.shape[3]
    yd = torch.empty(1, ph, pw, c, device="cuda")
    xd = nhwc(x).cuda()
    _lib.check(lib.maua_pool3x3_fwd(_lib.ptr(xd), _lib.ptr(yd), 1, h, w, c, avg, _lib.stream_ptr()), "pool3 fwd")
    torch.cuda.synchronize()
    assert rel(nchw(yd), y.detach()) < 1e-6
    gy = torch.randn(1, c, ph, pw, generator=g)
    add = torch.randn(1, c, h, w, generator=g)
    (gref,) = torch.autograd.grad(y, xr, gy)
    gref = gref + add * (x > 0)
    gx = torch.empty(1, h, w, c, device="cuda")
    gyd, addd = nhwc(gy).cuda(), nhwc(add).cuda()
    _lib.check(lib.maua_pool3x3_bwd(_lib.ptr(xd), _lib.ptr(gyd), _lib.ptr(addd), _lib.ptr(gx), 1, h, w, c, avg,
                                    0, _lib.stream_ptr()), "pool3 bwd")
    torch.cuda.synchronize()
    assert rel(nchw(gx), gref) < 1e-6


@pytest.mark.parametrize("impl", [pytest.param(_lib.MAUA_IMPL_REF, id="ref"), pytest.param(_lib.MAUA_IMPL_TC, id="tc"),
                                  pytest.param(_lib.MAUA_IMPL_TC_1CTA, id="tc1cta"), pytest.param(_lib.MAUA_IMPL_TC_2CTA, id="tc2cta")])
@pytest.mark.parametrize("cin,cout,ks,h,w", [(128, 256, 5, 40, 56), (96 + 32, 64, 5, 17, 33), (256, 128, 5, 64, 64), (64, 128, 5, 9, 5),
                                              (96 + 32, 128, 1, 31, 45), (384, 384, 1, 12, 20)])
def test_conv_kxk_tensor_core(lib, impl, cin, cout, ks, h, w):
    """NIN's 5x5 / pad 2 and 1x1 layers through conv_tc_kernel (KS = 5 halo pipeline / pointwise), forward and -- with the
    rotated, transposed weights -- input gradient, against fp32 torch on TF32-rounded operands."""
    g = torch.Generator().manual_seed(cin + cout + h)
    x = tf32_round(torch.randn(1, cin, h, w, generator=g))
    wt = torch.randn(cout, cin, ks, ks, generator=g) * (2.0 / (cin * ks * ks)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    xr = x.double().requires_grad_(True)
    pre = F.conv2d(xr, tf32_round(wt).double(), b.double(), padding=ks // 2)
    ref = F.relu(pre).float()
    wd_, bd, xd = wt.cuda(), b.cuda(), nhwc(x).cuda()
    wg = torch.empty(cout, ks * ks * cin, device="cuda")
    _lib.check(lib.maua_prep_conv_weights_k(_lib.ptr(wd_), _lib.ptr(wg), cout, cin, ks, 0, 1, _lib.stream_ptr()))
    y = torch.empty(1, h, w, cout, device="cuda")
    _lib.check(lib.maua_conv_kxk_fwd(_lib.ptr(xd), _lib.ptr(wg), _lib.ptr(bd), _lib.ptr(y), 1, h, w, cin, cout, ks, 1, impl,
                                     _lib.stream_ptr()), "conv_kxk_fwd")
    torch.cuda.synchronize()
    assert rel(nchw(y), ref) < 1.5e-3
    # input gradient: the same kernel on the rotated / transposed weights
    go = tf32_round(torch.randn(1, cout, h, w, generator=g))
    (gref,) = torch.autograd.grad(pre, xr, go.double())
    wdg = torch.empty(cin, ks * ks * cout, device="cuda")
    _lib.check(lib.maua_prep_conv_weights_k(_lib.ptr(wd_), _lib.ptr(wdg), cout, cin, ks, 1, 1, _lib.stream_ptr()))
    god = nhwc(go).cuda()
    gx = torch.empty(1, h, w, cin, device="cuda")
    _lib.check(lib.maua_conv_kxk_fwd(_lib.ptr(god), _lib.ptr(wdg), C.c_void_p(0), _lib.ptr(gx), 1, h, w, cout, cin, ks, 0, impl,
                                     _lib.stream_ptr()), "conv_kxk dgrad")
    torch.cuda.synchronize()
    assert rel(nchw(gx), gref.float()) < 1.5e-3
