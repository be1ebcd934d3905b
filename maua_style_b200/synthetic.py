"""Seeded synthetic inputs, random-init VGG-19 checkpoints and reference-shaped args for benchmarks and smoke runs.

Pretrained checkpoints and image files are unavailable offline, so throughput is measured on synthetic data of the
reference's value range: images are `U(0,255) - BGR mean` (what load.preprocess produces, load.py:21-32) and weights are
He-normal (default conv init makes deep VGG features vanish, SURVEY.md section 8a hazard 4).  The generators are
bit-identical to the ones the CPU oracle uses for its golden vectors (tests/test_oracle.py checks that), but this module
is product code and does not import the oracle.
"""
from __future__ import annotations

import argparse
import json
import math
from pathlib import Path

import torch
import torch.nn.functional as F

from .models import NIN_FEATURE_INDEX, NIN_LAYERS, channel_list

BGR_MEAN = (103.939, 116.779, 123.68)  # load.py:30


def synthetic_image(h: int, w: int, seed: int, smooth: bool = False) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    if smooth:  # bilinear-upsampled x8 noise: non-trivial ReLU sparsity / spatial correlation
        low = torch.rand(1, 3, max(h // 8, 2), max(w // 8, 2), generator=g)
        img = F.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)
    else:
        img = torch.rand(1, 3, h, w, generator=g)
    return img * 255.0 - torch.tensor(BGR_MEAN).view(1, 3, 1, 1)


def he_init(seed: int = 0, channels=None):
    """[(weight [Cout,Cin,3,3], bias [Cout])] per conv: He-normal weights, bias sigma 0.1."""
    channels = channels or channel_list["VGG-19"]
    g = torch.Generator().manual_seed(seed)
    params, cin = [], 3
    for c in channels:
        if c == "P":
            continue
        c, ks = c if isinstance(c, tuple) else (c, 3)  # NIN_LAYERS entries are (channels, kernel size)
        w = torch.randn(c, cin, ks, ks, generator=g) * math.sqrt(2.0 / (cin * ks * ks))
        b = torch.randn(c, generator=g) * 0.1
        params.append((w, b))
        cin = c
    return params


def save_random_checkpoint(path, seed: int = 0, channels=None):
    """torchvision-style state dict (`features.N.weight / bias`, models.py:343-347) with the seeded weights."""
    channels = channels or channel_list["VGG-19"]
    params = he_init(seed, channels)
    sd, k, ci = {}, 0, 0
    for c in channels:
        if c == "P":
            k += 1
            continue
        if channels is NIN_LAYERS:
            k = NIN_FEATURE_INDEX[ci]  # position of the conv inside the reference's NIN.features (models.py:82-111)
        w, b = params[ci]
        sd[f"features.{k}.weight"] = w.clone()
        sd[f"features.{k}.bias"] = b.clone()
        ci += 1
        k += 2
    torch.save(sd, path)
    return params


def reference_args(ckpt_path, workdir, **over) -> argparse.Namespace:
    """An args Namespace with the fields config.get_args + config.postprocess produce (config.py:12-168), CLI defaults."""
    a = argparse.Namespace(
        transfer_type="img_img", model_file=str(ckpt_path), pooling="max", disable_check=True,
        content_layers="relu4_2", style_layers="relu1_1,relu2_1,relu3_1,relu4_1,relu5_1",
        content_weight=5.0, style_weight=100.0, tv_weight=1e-3, temporal_weight=50.0,
        use_covariance=False, normalize_gradients=True, normalize_weights=False, video_style_factor=100.0,
        shift_factor=0, style_blend_weights=[1.0], optimizer="adam", learning_rate=1.0, lbfgs_num_correction=100,
        lbfgs_tolerance_change=-1, lbfgs_tolerance_grad=-1, gpu="0", multidevice=False, multidevice_strategy="5",
        verbose=False, print_iter=0, save_iter=0, num_iters=[10], image_sizes=[64], backward_device="cuda:0",
    )
    over = dict(over)
    if "style_blend_weights" in over and isinstance(over["style_blend_weights"], str):
        w = [float(x) for x in over.pop("style_blend_weights").split(",")]
        a.style_blend_weights = [x / sum(w) for x in w]  # config.py:160-164
    for k, v in over.items():
        setattr(a, k, v)
    scaling = Path(workdir) / "scaling.json"
    scaling.write_text(json.dumps({"100000": {"model_file": str(ckpt_path), "optimizer": a.optimizer,
                                              "multidevice": bool(a.multidevice), "gpu": a.gpu}}))
    a.scaling_args = str(scaling)
    return a
