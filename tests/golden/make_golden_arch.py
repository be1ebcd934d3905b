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
