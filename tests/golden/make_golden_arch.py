"""Golden vectors for the architecture / layer-list variants of the hot path, produced by the UNMODIFIED reference on the
CPU (same recipe as make_golden.py, whose helpers this script re-uses):

    python tests/golden/make_golden_arch.py

  vgg16_adam_gram_72x88 ........ `--model_file *vgg16*` (models.py:334-347: VGG-16 channel list, vgg16_dict names), default
                                 layers, 3 Adam iterations
  vid_frame_temporal_64x80 ..... one vid_img frame (style.py:276-294): temporal target + flow-reliability weights captured
                                 with optim.set_temporal_targets, then optimize(content, styles, init, n, args, net, losses)
  vgg19_deep_taps_avg_64x96 .... two content taps (relu3_2, relu5_2) and style taps relu2_2 / relu4_3: the stack runs one
                                 layer past relu5_1, average pooling, 3 Adam iterations
  vgg19_same_layer_taps_64x64 .. ContentLoss and StyleLoss on the same ReLU (relu4_2; models.py:411-431 inserts the content
                                 module first)
  vgg16p_adam_gram_72x88 ....... `--model_file *prun*` (models.py:249-258: the channel-pruned VGG-16, channel list "VGG-16p" =
                                 24, 22, 41, 51, 108, 89, 111, 184, 276, 228, 512...), default layers, 3 Adam iterations
  vgg16p_cov_lbfgs_64x80 ....... the same stack with the covariance loss, 2 blended styles, 4 L-BFGS iterations
  vgg19_taps_lbfgs_80x64 ....... VGG-19 with `--style_layers relu1_2,relu3_3 --content_layers relu2_2`: taps that sit
                                 directly before a pool, truncation after relu3_3 (models.py:382), 4 L-BFGS iterations
"""
from __future__ import annotations

import os
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as mg  # noqa: E402

O = mg.O


def video_inputs(T, h, w, style_shapes):
    """Seeded stand-ins for an img_vid job: content image, style clips [(frames, h, w)], pastiche video of T frames."""
    import torch

    content = O.synthetic_image(h, w, seed=1, smooth=True)
    styles = [torch.cat([O.synthetic_image(sh, sw, seed=20 + 10 * i + f, smooth=(f % 2 == 0)) for f in range(n)])
              for i, (n, sh, sw) in enumerate(style_shapes)]
    init = torch.cat([O.synthetic_image(h, w, seed=40 + f) * 0.25 for f in range(T)])
    return content, styles, init


def run_video_case(name, rconfig, rmodels, roptim, workdir, ckpt, T, h, w, style_shapes, gfw, afw, iters, **over):
    """transfer_type img_vid (optim.py:113-125, :149-170, :216-219): one feval on the first window of `gfw` frames (per-module
    losses, the static / dynamic style targets, the gradient of every frame) and the whole windowed optimisation."""
    import json

    import numpy as np
    import torch

    torch.manual_seed(0)
    torch.set_flush_denormal(True)
    args = mg.reference_args(rconfig, workdir, ckpt, n_styles=len(style_shapes), **over)
    args.transfer_type, args.gram_frame_window, args.avg_frame_window = "img_vid", gfw, afw
    content, styles, init = video_inputs(T, h, w, style_shapes)
    out = {"meta": json.dumps(dict(name=name, T=T, h=h, w=w, style_shapes=style_shapes, gfw=gfw, afw=afw, iters=iters, over=over,
                                   blend=[float(x) for x in args.style_blend_weights], arch="VGG-19"))}
    net, losses = rmodels.load_model(args)
    roptim.set_content_targets(net, content, args)
    first = styles if afw == -1 else [s[:afw] if s.shape[0] > 1 else s for s in styles]
    roptim.set_style_video_targets(net, first, args)
    for m in losses:
        m.mode = "loss"
    x = init[:gfw].clone().requires_grad_(True)
    net(x)
    total = 0
    for i, m in enumerate(losses):
        if isinstance(m.loss, int):
            out[f"loss_{i}_{m.name.split()[0]}"] = np.float32(0)
            continue
        out[f"loss_{i}_{m.name.split()[0]}"] = np.float32(m.loss.item())
        total = total + m.loss
    total.backward()
    out["grad"] = x.grad.numpy().astype(np.float32)
    out["total"] = np.float32(total.item())
    for i, m in enumerate(net.style_losses):
        out[f"style_target_{i}_stats"] = np.array([m.target.sum().item(), m.target.norm().item()], dtype=np.float64)
        out[f"style_target_{i}_block"] = m.target[:16, :16].numpy().astype(np.float32)
        vt = m.video_target
        out[f"video_target_{i}_shape"] = np.array(vt.shape, dtype=np.int64)
        out[f"video_target_{i}_stats"] = np.array([vt.sum().item(), vt.norm().item()], dtype=np.float64)
        out[f"video_target_{i}_sample"] = mg.sample(vt)
    for m in losses:
        m.loss = 0
    res = roptim.optimize(content, styles, init.clone(), iters, args)
    out["optimized"] = res.detach().numpy().astype(np.float32)
    np.savez_compressed(mg.HERE / f"{name}.npz", **out)
    print(f"wrote {name}.npz ({(mg.HERE / (name + '.npz')).stat().st_size / 1024:.0f} KiB)")


def save_nin_checkpoint(rmodels, path, seed=0):
    """State dict of the reference's NIN module (models.py:74-113) with the oracle's seeded He-normal weights in its 12 convs."""
    import torch

    params = O.he_init_vgg19(seed, O.NIN_LAYERS)
    net = rmodels.NIN("max")
    sd = net.state_dict()
    keys = [k for k in sd if k.endswith(".weight")]
    assert len(keys) == len(params) == 12
    for (w, b), k in zip(params, keys):
        assert sd[k].shape == w.shape, (k, sd[k].shape, w.shape)
        sd[k] = w.clone()
        sd[k.replace(".weight", ".bias")] = b.clone()
    torch.save(sd, path)
    return params


def main():
    only = set(sys.argv[1:])  # optional: names of the cases to (re)generate
    rconfig, rloss, rmodels, roptim = mg.import_reference()
    run = mg.run_case
    if only:
        mg.run_case = lambda name, **kw: run(name, **kw) if name in only else None
    with tempfile.TemporaryDirectory() as td:
        workdir = Path(td)
        os.chdir(workdir)
        ckpt16p = workdir / "vgg16-prune-random.pth"
        mg.save_checkpoint(rmodels, ckpt16p, arch="VGG-16p")
        names16 = O.relu_names(O.VGG16P_CHANNELS)
        pre = dict(rconfig=rconfig, rmodels=rmodels, roptim=roptim, workdir=workdir)
        mg.run_case("vgg16p_adam_gram_72x88", ckpt=ckpt16p, h=72, w=88, style_hw=[(64, 96)], iters=3, relu_names=names16,
                    meta_extra={"arch": "VGG-16p"}, **pre)
        mg.run_case("vgg16p_cov_lbfgs_64x80", ckpt=ckpt16p, h=64, w=80, style_hw=[(72, 72), (56, 96)], iters=4,
                    optimizer="lbfgs", use_covariance=True, style_blend_weights="3,1", relu_names=names16,
                    meta_extra={"arch": "VGG-16p"}, **pre)
        ckpt19v = workdir / "vgg19-random.pth"
        mg.save_checkpoint(rmodels, ckpt19v)
        if not only or "img_vid_windows_adam_48x64" in only:
            # a style video (4 frames) and a style image, all targets captured once (avg_frame_window -1), 3 windows of 3 frames
            run_video_case("img_vid_windows_adam_48x64", ckpt=ckpt19v, T=5, h=48, w=64, style_shapes=[(4, 56, 56), (1, 48, 48)],
                           gfw=3, afw=-1, iters=2, style_blend_weights="3,1", **pre)
        if not only or "img_vid_avgwin_lbfgs_48x48" in only:
            # style targets re-captured per window from 2-frame style windows: the dynamic target is [2C, 2C] and the dynamic
            # term of the 3-frame pastiche windows is skipped (loss.py:165-166); covariance loss; L-BFGS
            run_video_case("img_vid_avgwin_lbfgs_48x48", ckpt=ckpt19v, T=4, h=48, w=48, style_shapes=[(3, 48, 64)],
                           gfw=3, afw=2, iters=3, optimizer="lbfgs", use_covariance=True, **pre)
        ckptnin = workdir / "nin-random.pth"
        save_nin_checkpoint(rmodels, ckptnin)
        nin_names = O.relu_names(O.NIN_LAYERS)
        # the layer lists the stock config/scaling-img.json uses for NIN (:32-47)
        mg.run_case("nin_adam_gram_131x150", ckpt=ckptnin, h=131, w=150, style_hw=[(140, 128)], iters=3, relu_names=nin_names,
                    style_layers="relu1,relu3,relu5,relu7,relu9,relu11", content_layers="relu8", meta_extra={"arch": "NIN"}, **pre)
        # taps in front of the pools, the 1000-channel cccp8 layer as content tap, average pooling, covariance loss, L-BFGS
        mg.run_case("nin_avg_cov_lbfgs_128x144", ckpt=ckptnin, h=128, w=144, style_hw=[(120, 150), (136, 130)], iters=4,
                    optimizer="lbfgs", pooling="avg", use_covariance=True, style_blend_weights="3,1", relu_names=nin_names,
                    style_layers="relu3,relu6,relu10", content_layers="relu12", meta_extra={"arch": "NIN"}, **pre)
        ckpt16 = workdir / "vgg16-random.pth"
        mg.save_checkpoint(rmodels, ckpt16, arch="VGG-16")
        common = dict(rconfig=rconfig, rmodels=rmodels, roptim=roptim, workdir=workdir)
        mg.run_case("vgg16_adam_gram_72x88", ckpt=ckpt16, h=72, w=88, style_hw=[(64, 96)], iters=3,
                    relu_names=O.relu_names(O.VGG16_CHANNELS), meta_extra={"arch": "VGG-16"}, **common)
        ckpt19 = workdir / "vgg19-random.pth"
        mg.save_checkpoint(rmodels, ckpt19)
        mg.run_case("vid_frame_temporal_64x80", ckpt=ckpt19, h=64, w=80, style_hw=[(72, 72)], iters=4, temporal=True,
                    meta_extra={"arch": "VGG-19", "temporal": True}, **common)
        mg.run_case("vgg19_deep_taps_avg_64x96", ckpt=ckpt19, h=64, w=96, style_hw=[(80, 80)], iters=3,
                    style_layers="relu2_2,relu4_3", content_layers="relu3_2,relu5_2", pooling="avg",
                    meta_extra={"arch": "VGG-19"}, **common)
        mg.run_case("vgg19_same_layer_taps_64x64", ckpt=ckpt19, h=64, w=64, style_hw=[(64, 64)], iters=3,
                    style_layers="relu1_1,relu4_2", content_layers="relu4_2", meta_extra={"arch": "VGG-19"}, **common)
        mg.run_case("vgg19_taps_lbfgs_80x64", ckpt=ckpt19, h=80, w=64, style_hw=[(72, 72)], iters=4, optimizer="lbfgs",
                    style_layers="relu1_2,relu3_3", content_layers="relu2_2", meta_extra={"arch": "VGG-19"}, **common)


if __name__ == "__main__":
    main()
