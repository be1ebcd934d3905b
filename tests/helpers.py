"""Shared helpers of the GPU parity tests: reference-style args, seeded checkpoint, golden loading."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import maua_oracle as O  # noqa: E402  (tests may use the oracle; the product never does)

GOLDEN = ROOT / "tests" / "golden"


def make_args(ckpt_path, tmpdir, **over):
    """An args Namespace with the fields the reference's config.postprocess produces (config.py:134-168)."""
    a = argparse.Namespace(
        transfer_type="img_img", model_file=str(ckpt_path), pooling="max", disable_check=True,
        content_layers="relu4_2", style_layers="relu1_1,relu2_1,relu3_1,relu4_1,relu5_1",
        content_weight=5.0, style_weight=100.0, tv_weight=1e-3, temporal_weight=50.0,
        use_covariance=False, normalize_gradients=True, normalize_weights=False, video_style_factor=100.0,
        shift_factor=0, style_blend_weights=[1.0], optimizer="adam", learning_rate=1.0, lbfgs_num_correction=100,
        lbfgs_tolerance_change=-1, lbfgs_tolerance_grad=-1, gpu="0", multidevice=False, multidevice_strategy="5",
        verbose=False, print_iter=0, save_iter=0, num_iters=[10], image_sizes=[64], backward_device="cuda:0",
    )
    scaling = Path(tmpdir) / "scaling.json"
    over = dict(over)
    if "no_grad_norm" in over:
        a.normalize_gradients = not over.pop("no_grad_norm")
    if "style_blend_weights" in over and isinstance(over["style_blend_weights"], str):
        w = [float(x) for x in over.pop("style_blend_weights").split(",")]
        a.style_blend_weights = [x / sum(w) for x in w]
    for k, v in over.items():
        setattr(a, k, v)
    scaling.write_text(json.dumps({"100000": {"model_file": str(ckpt_path), "optimizer": a.optimizer,
                                              "multidevice": False, "gpu": a.gpu}}))
    a.scaling_args = str(scaling)
    return a


def save_checkpoint(path, seed=0, channels=O.VGG19_CHANNELS):
    """torchvision-style state dict (features.N.weight / bias) with the seeded He-normal VGG-19 (or VGG-16) weights."""
    params = O.he_init_vgg19(seed, channels)
    sd, k, ci = {}, 0, 0
    nin_index = [0, 2, 4, 7, 9, 11, 14, 16, 18, 22, 24, 26]  # convs inside the reference's NIN.features (models.py:82-111)
    for c in channels:
        if c == "P":
            k += 1
            continue
        if O.is_nin(channels):
            k = nin_index[ci]
        w, b = params[ci]
        sd[f"features.{k}.weight"] = w.clone()
        sd[f"features.{k}.bias"] = b.clone()
        ci += 1
        k += 2
    torch.save(sd, path)
    return params


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    return z, json.loads(str(z["meta"]))


def golden_inputs(meta):
    content = O.synthetic_image(meta["h"], meta["w"], seed=1, smooth=True)
    styles = [O.synthetic_image(sh, sw, seed=2 + i, smooth=(i % 2 == 1)) for i, (sh, sw) in enumerate(meta["style_hw"])]
    init = O.synthetic_image(meta["h"], meta["w"], seed=4) * 0.25
    return content, styles, init


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def video_inputs(meta):
    """make_golden_arch.video_inputs: content image, style clips, pastiche video of an img_vid golden."""
    content = O.synthetic_image(meta["h"], meta["w"], seed=1, smooth=True)
    styles = [torch.cat([O.synthetic_image(sh, sw, seed=20 + 10 * i + f, smooth=(f % 2 == 0)) for f in range(n)])
              for i, (n, sh, sw) in enumerate(meta["style_shapes"])]
    init = torch.cat([O.synthetic_image(meta["h"], meta["w"], seed=40 + f) * 0.25 for f in range(meta["T"])])
    return content, styles, init


def video_cfg(meta):
    over = dict(meta["over"])
    cfg = O.StyleConfig(content_weight=5.0)
    cfg.optimizer = over.pop("optimizer", "adam")
    if "style_blend_weights" in over:
        cfg.style_blend_weights = [float(x) for x in over.pop("style_blend_weights").split(",")]
    for k, v in over.items():
        assert hasattr(cfg, k), k
        setattr(cfg, k, v)
    return cfg
