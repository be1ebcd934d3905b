"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference (JCBrouwer/maua-style,
mounted read-only at /root/reference or $MAUA_REF) on seeded synthetic inputs, on the CPU.

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as small .npz fixtures; this script is
committed next to them so they can be regenerated.  The reference has no tests or fixtures of its own for this
path, so these differential vectors are what pins the oracle (oracle/maua_oracle.py) and,
through it, the CUDA path.

Recipe: stub the optional imports the hot path never touches (gdown, skvideo, ffmpeg,
flow), import loss / models / optim / config from the reference, save a He-normal random-init VGG-19
checkpoint (pretrained weights are unavailable offline) under a path containing "vgg19" and load it with
disable_check=True (strict=False, models.py:343), then drive optim.optimize exactly as style.py:69 does.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
REF = Path(os.environ.get("MAUA_REF", "/root/reference"))

from oracle import maua_oracle as O  # noqa: E402  (only for the shared seeded weight / input generators)


def import_reference():
    for name in ("gdown", "skvideo", "skvideo.io", "ffmpeg", "flow"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["skvideo"].io = sys.modules["skvideo.io"]
    sys.path.insert(0, str(REF))
    import config as rconfig  # noqa
    import loss as rloss  # noqa
    import models as rmodels  # noqa
    import optim as roptim  # noqa

    return rconfig, rloss, rmodels, roptim


def reference_args(rconfig, workdir: Path, ckpt: Path, **over):
    args = argparse.Namespace()
    with open(REF / "config" / "args-img.json") as f:
        args.__dict__ = json.load(f)
    scaling = workdir / "scaling.json"
    optimizer = over.pop("optimizer", "adam")
    scaling.write_text(json.dumps({"100000": {"model_file": str(ckpt), "optimizer": optimizer, "multidevice": False, "gpu": "c"}}))
    args.__dict__.update(dict(gpu="c", backend="mkl", scaling_args=str(scaling), model_file=str(ckpt), disable_check=True,
                              no_hist_match=True, content_weight=5.0, optimizer=optimizer, style=["s"] * over.pop("n_styles", 1),
                              content="c", output_dir=str(workdir)))
    args.__dict__.update(over)
    args = rconfig.postprocess(args)
    return args


def save_checkpoint(rmodels, path: Path, seed: int = 0, arch: str = "VGG-19"):
    channels = {"VGG-19": O.VGG19_CHANNELS, "VGG-16": O.VGG16_CHANNELS, "VGG-16p": O.VGG16P_CHANNELS}[arch]
    params = O.he_init_vgg19(seed, channels)
    seq = rmodels.build_sequential(rmodels.channel_list[arch], "max")
    sd = seq.state_dict()
    keys = [k for k in sd if k.endswith(".weight")]
    assert len(keys) == len(params)
    # only the 13 convs up to conv5_1 matter (the net is truncated at relu5_1); the rest keep their default init
    for (w, b), k in zip(params, keys):
        sd[k] = w.clone()
        sd[k.replace(".weight", ".bias")] = b.clone()
    torch.save({f"features.{k}": v for k, v in sd.items()}, path)
    return params


def sample(t: torch.Tensor, n: int = 512) -> np.ndarray:
    """Deterministic strided sample of a tensor (keeps the fixtures small)."""
    flat = t.detach().reshape(-1)
    idx = torch.linspace(0, flat.numel() - 1, min(n, flat.numel())).long()
    return flat[idx].numpy().astype(np.float32)


def temporal_inputs(h, w):
    """Stand-ins for the warped previous frame and the flow-reliability map of vid_img (style.py:279-288)."""
    warp = O.synthetic_image(h, w, seed=9, smooth=True)
    weights = torch.rand(1, 1, h, w, generator=torch.Generator().manual_seed(3))
    return warp, weights


def run_case(name, rconfig, rmodels, roptim, workdir, ckpt, h, w, style_hw, iters, relu_names=None, meta_extra=None,
             temporal=False, **over):
    torch.manual_seed(0)
    torch.set_flush_denormal(True)
    n_styles = len(style_hw)
    args = reference_args(rconfig, workdir, ckpt, n_styles=n_styles, **over)
    content = O.synthetic_image(h, w, seed=1, smooth=True)
    styles = [O.synthetic_image(sh, sw, seed=2 + i, smooth=(i % 2 == 1)) for i, (sh, sw) in enumerate(style_hw)]
    init = O.synthetic_image(h, w, seed=4) * 0.25

    relu_names = relu_names or O.VGG19_RELU_NAMES
    out = {"meta": json.dumps(dict(name=name, h=h, w=w, style_hw=style_hw, iters=iters, over=over,
                                   blend=[float(x) for x in args.style_blend_weights], **(meta_extra or {})))}
    # --- one feval: per-module losses, targets, image gradient -------------------------------------------
    net, losses = rmodels.load_model(args)
    if temporal:  # style.py:288 (vid_img): the temporal target is captured before optimize() is called with net / losses
        roptim.set_temporal_targets(net, *temporal_inputs(h, w), args=args)
    roptim.set_content_targets(net, content, args)
    roptim.set_style_targets(net, styles, args)
    for m in losses:
        m.mode = "loss"
    taps = {}
    hooks = []
    relu_i = 0
    import torch.nn as nn
    for mod in net:
        if isinstance(mod, nn.ReLU):
            nm = relu_names[relu_i]
            relu_i += 1
            hooks.append(mod.register_forward_hook(lambda m, i, o, nm=nm: taps.__setitem__(nm, o.detach().clone())))
    x = init.clone().requires_grad_(True)
    net(x)
    for hnd in hooks:
        hnd.remove()
    total = 0
    for i, m in enumerate(losses):
        if isinstance(m.loss, int):  # module skipped (e.g. temporal loss without a target): optim.py:208-209
            out[f"loss_{i}_{m.name.split()[0]}"] = np.float32(0)
            continue
        out[f"loss_{i}_{m.name.split()[0]}"] = np.float32(m.loss.item())
        total = total + m.loss
    total.backward()
    out["grad"] = x.grad.numpy().astype(np.float32)
    out["total"] = np.float32(total.item())
    for nm, t in taps.items():
        out[f"feat_{nm}_stats"] = np.array([t.sum().item(), t.norm().item(), t.abs().max().item()], dtype=np.float64)
        out[f"feat_{nm}_sample"] = sample(t)
    for i, m in enumerate(net.style_losses):
        out[f"style_target_{i}_stats"] = np.array([m.target.sum().item(), m.target.norm().item()], dtype=np.float64)
        out[f"style_target_{i}_sample"] = sample(m.target)
        out[f"style_target_{i}_block"] = m.target[:16, :16].numpy().astype(np.float32)
    for m in losses:
        m.loss = 0
    # --- N iterations through optim.optimize (fresh model, as style.py:69 does) --------------------------
    if iters > 0 and temporal:  # style.py:176-177, :288-294: load_model once, temporal target, optimize(..., net, losses)
        net2, losses2 = rmodels.load_model(args)
        roptim.set_temporal_targets(net2, *temporal_inputs(h, w), args=args)
        res = roptim.optimize(content, styles, init.clone(), iters, args, net2, losses2)
        out["optimized"] = res.detach().numpy().astype(np.float32)
    elif iters > 0:
        res = roptim.optimize(content, styles, init.clone(), iters, args)
        out["optimized"] = res.detach().numpy().astype(np.float32)
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(f"wrote {name}.npz ({(HERE / (name + '.npz')).stat().st_size / 1024:.0f} KiB)")


def main():
    rconfig, rloss, rmodels, roptim = import_reference()
    with tempfile.TemporaryDirectory() as td:
        workdir = Path(td)
        os.chdir(workdir)
        ckpt = workdir / "vgg19-random.pth"
        save_checkpoint(rmodels, ckpt)
        common = dict(rconfig=rconfig, rmodels=rmodels, roptim=roptim, workdir=workdir, ckpt=ckpt)
        # config 1 family: Gram loss, 1 style, Adam (10 iters = 11 evals) at 64^2 and an odd size
        run_case("adam_gram_64", h=64, w=64, style_hw=[(64, 64)], iters=10, **common)
        run_case("adam_gram_90x122", h=90, w=122, style_hw=[(101, 75)], iters=5, **common)
        # config 2 family: L-BFGS, content + style + TV
        run_case("lbfgs_gram_64", h=64, w=64, style_hw=[(64, 64)], iters=10, optimizer="lbfgs", **common)
        # config 3 family: covariance loss, 2 blended styles (3:1) of different shapes
        run_case("adam_cov_2styles_96x128", h=96, w=128, style_hw=[(80, 120), (128, 96)], iters=5, use_covariance=True,
                 style_blend_weights="3,1", **common)
        # switches: no gradient normalisation, no dynamic term, avg pooling, normalize_weights
        run_case("adam_nonorm_novsf_avg_64", h=64, w=64, style_hw=[(64, 64)], iters=3, no_grad_norm=True, video_style_factor=0,
                 pooling="avg", **common)
        # (normalize_weights + the default temporal module divides by max(empty size) = 0 in the reference, optim.py:178,
        #  so this case runs without the temporal module, as the reference itself requires)
        run_case("adam_normweights_notemporal_64", h=64, w=64, style_hw=[(48, 80)], iters=3, normalize_weights=True,
                 temporal_weight=0, **common)


if __name__ == "__main__":
    main()
