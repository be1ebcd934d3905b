"""Layer-wise multidevice split (maua_style_b200/parallel.py; reference models.py:503-566).

CPU: the module-index -> stage-boundary mapping.  GPU: a staged plan must reproduce the single plan (same kernels, the
hand-over only changes where one tensor lives); with one visible GPU both stages are placed on cuda:0, with two or more
the second stage runs on cuda:1 and the hand-over goes through NVLink peer stores.
"""
import pytest
import torch

from maua_style_b200 import parallel

ENTRIES = [64, 64, 0, 128, 128, 0, 256, 256, 256, 256, 0, 512, 512, 512, 512, 0, 512]  # VGG-19 up to relu5_1
TAPS = [(0, "style"), (2, "style"), (4, "style"), (8, "style"), (9, "content"), (12, "style")]


def test_module_index_map_matches_the_reference_sequential():
    # SURVEY.md section 3.3 (printed reference net, default args): relu1_1 = 3, style = 4, relu1_2 = 6, pool1 = 7,
    # style after relu2_1 = 10, pool2 = 13, ..., content after relu4_2 = 29, pool4 = 34, style after relu5_1 = 37
    last = parallel.module_index_map(ENTRIES, TAPS, True, True)
    assert last == [4, 6, 7, 10, 12, 13, 16, 18, 20, 22, 23, 26, 29, 31, 33, 34, 37]
    assert parallel.module_index_map(ENTRIES, TAPS, False, False)[0] == 2


def test_stage_bounds():
    # reference default "5" = after conv1_2 -> moved to the end of that entry and past pool1
    assert parallel.stage_bounds(ENTRIES, "5", 2, TAPS, True, True) == [0, 3, 17]
    assert parallel.stage_bounds(ENTRIES, "13,23", 3, TAPS, True, True) == [0, 6, 11, 17]
    assert parallel.stage_bounds(ENTRIES, "3", 2, TAPS, True, True) == [0, 1, 17]  # after relu1_1 (+ its style module)
    with pytest.raises(AssertionError):
        parallel.stage_bounds(ENTRIES, "5,9", 2, TAPS, True, True)  # models.py:541-543
    with pytest.raises(ValueError):
        parallel.stage_bounds(ENTRIES, "37", 2, TAPS, True, True)  # nothing left for the second device


def _feval(net, losses, args, content, styles, init):
    from maua_style_b200 import optim

    optim.set_content_targets(net, content, args)
    optim.set_style_targets(net, styles, args)
    for m in losses:
        m.mode = "loss"
    vec, g = optim.feval(net, init.clone().to(net.device))
    return vec.clone().cpu(), g.clone().cpu()


@pytest.mark.gpu
@pytest.mark.parametrize("strategy,ngpu", [("5", 2), ("3", 2), ("13,29", 3), ("22", 2)])
@pytest.mark.parametrize("pooling", ["max", "avg"])
def test_staged_plan_matches_single_plan(tmp_path, strategy, ngpu, pooling):
    from helpers import O, make_args, rel, save_checkpoint
    from maua_style_b200 import models

    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    content = O.synthetic_image(72, 104, seed=1, smooth=True)
    styles = [O.synthetic_image(80, 96, seed=2)]
    init = O.synthetic_image(72, 104, seed=4) * 0.25
    a1 = make_args(ckpt, tmp_path, pooling=pooling)
    net1, losses1 = models.load_model(a1)
    vec1, g1 = _feval(net1, losses1, a1, content, styles, init)

    have = torch.cuda.device_count()
    gpus = ",".join(str(i % have) for i in range(ngpu))  # fewer GPUs than stages: stages share devices
    a2 = make_args(ckpt, tmp_path, pooling=pooling, multidevice=True, gpu=gpus, multidevice_strategy=strategy)
    net2, losses2 = models.load_model(a2)
    assert net2.n_stages == ngpu
    vec2, g2 = _feval(net2, losses2, a2, content, styles, init)
    assert torch.allclose(vec1, vec2, rtol=1e-5, atol=0), (vec1, vec2)
    err = rel(g2, g1)
    print(f"staged ({strategy}, {gpus}, {pooling}) vs single plan: gradient rel {err:.2e}")
    # identical kernels; only the fp32 summation order of "dgrad + tap gradient" differs at a boundary entry
    assert err < (1e-4 if pooling == "avg" else 2e-3)

    # the reference-shaped autograd interface works across stages too
    x = init.clone().to(net2.device).requires_grad_(True)
    net2(x)
    total = sum(m.loss for m in losses2 if not isinstance(m.loss, int))
    total.backward()
    assert rel(x.grad.cpu(), g2) < 1e-6
    for m in losses2:
        m.loss = 0


@pytest.mark.gpu
def test_staged_optimize_runs(tmp_path):
    from helpers import O, make_args, save_checkpoint
    from maua_style_b200 import models, optim

    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    have = torch.cuda.device_count()
    content = O.synthetic_image(64, 64, seed=1, smooth=True)
    styles = [O.synthetic_image(64, 64, seed=2)]
    init = O.synthetic_image(64, 64, seed=4) * 0.25
    outs = []
    for multi in (False, True):
        a = make_args(ckpt, tmp_path, multidevice=multi, gpu=("0," + str(1 % have)) if multi else "0",
                      multidevice_strategy="5")
        net, losses = models.load_model(a)
        outs.append(optim.optimize(content, styles, init.clone(), 5, a, net, losses))
    p = O.psnr(outs[0], outs[1])
    print(f"staged vs single optimize, 5 Adam iterations: PSNR {p:.1f} dB")
    assert p > 50.0


@pytest.mark.multigpu
@pytest.mark.parametrize("strategy", ["5", "13"])
def test_split_over_two_physical_gpus(tmp_path, strategy):
    """BASELINE.json configs[3] on real hardware (run with `gpurun --gpus 2 -- python -m pytest tests/test_parallel.py -m
    multigpu`): stage 0 on cuda:0, stage 1 on cuda:1, so the forward hand-over (dual TMA store from the conv / pool epilogue
    into the peer's memory), the backward hand-over (dgrad epilogue storing into cuda:0's g_top), cudaDeviceEnablePeerAccess,
    the cross-device events and the loss-vector gather all cross NVLink.  Against the single-GPU plan at 384 x 512."""
    from helpers import O, make_args, rel, save_checkpoint
    from maua_style_b200 import models, optim

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ckpt = tmp_path / "vgg19-random.pth"
    save_checkpoint(ckpt)
    h, w = 384, 512
    content = O.synthetic_image(h, w, seed=1, smooth=True)
    styles = [O.synthetic_image(h, w, seed=2)]
    init = O.synthetic_image(h, w, seed=4) * 0.25
    a1 = make_args(ckpt, tmp_path, pooling="avg")
    net1, losses1 = models.load_model(a1)
    vec1, g1 = _feval(net1, losses1, a1, content, styles, init)
    a2 = make_args(ckpt, tmp_path, pooling="avg", multidevice=True, gpu="0,1", multidevice_strategy=strategy)
    net2, losses2 = models.load_model(a2)
    devs = [str(st["device"]) for st in net2._stages]
    assert devs == ["cuda:0", "cuda:1"]
    vec2, g2 = _feval(net2, losses2, a2, content, styles, init)
    assert str(net2._stages[1]["x_in"].device) == "cuda:1" and str(net2._stages[0]["g_top"].device) == "cuda:0"
    assert torch.allclose(vec1, vec2, rtol=1e-5, atol=0), (vec1, vec2)
    err = rel(g2, g1)
    print(f"two physical GPUs, strategy {strategy}: gradient vs single-GPU plan rel {err:.2e}, "
          f"hand-over tensor {tuple(net2._stages[1]['x_in'].shape)} = {net2._stages[1]['x_in'].numel() * 4 / 1e6:.1f} MB each way")
    assert err < 1e-4
    # determinism across the link: a second evaluation is bit-identical
    vec3, g3 = _feval(net2, losses2, a2, content, styles, init)
    assert torch.equal(g2, g3) and torch.equal(vec2, vec3)
    # a whole optimisation (Adam, 6 iterations) over the two GPUs equals the single-GPU run
    outs = []
    for multi in (False, True):
        a = make_args(ckpt, tmp_path, pooling="avg", multidevice=multi, gpu="0,1" if multi else "0", multidevice_strategy=strategy)
        net, losses = models.load_model(a)
        outs.append(optim.optimize(content, styles, init.clone(), 6, a, net, losses))
    p = O.psnr(outs[0], outs[1])
    print(f"two physical GPUs, strategy {strategy}: 6 Adam iterations vs single GPU PSNR {p:.1f} dB")
    assert p > 60.0
