"""Bisect a gradient deviation of the exact-arithmetic plan against the fp64 oracle: by image extent and by which loss
modules are live.  Usage (GPU box): python tools/dbg_exact_bisect.py"""
import sys
import tempfile
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import O, make_args, rel, save_checkpoint  # noqa: E402

from maua_style_b200 import _lib, models, optim  # noqa: E402


def run(h, w, tmp, params, ckpt, impl=_lib.MAUA_IMPL_FP32, **over):
    args = make_args(ckpt, tmp, temporal_weight=0.0, **over)
    net, losses = models.load_model(args)
    net.set_impl(impl)
    content = O.synthetic_image(h, w, seed=1, smooth=True)
    style = O.synthetic_image(h, w, seed=2)
    init = O.synthetic_image(h, w, seed=4) * 0.25
    optim.set_content_targets(net, content, args)
    optim.set_style_targets(net, [style], args)
    for m in losses:
        m.mode = "loss"
    _, g = optim.feval(net, init.clone().cuda())
    cfg = O.StyleConfig(content_weight=args.content_weight, temporal_weight=0.0, tv_weight=args.tv_weight,
                        style_weight=args.style_weight, content_layers=args.content_layers, style_layers=args.style_layers,
                        pooling=args.pooling)
    onet = O.OracleNet([(a.double(), b.double()) for a, b in params], cfg)
    O.set_content_targets(onet, content.double())
    O.set_style_targets(onet, [style.double()], [1.0])
    for m in onet.losses:
        m.mode = "loss"
    _, _, g64 = O.feval(onet, init.double())
    d = (g.detach().cpu().double() - g64)[0].pow(2).sum(0)
    hh, ww = divmod(int(d.argmax()), d.shape[1])
    return rel(g, g64), (hh, ww)


def main():
    tmp = tempfile.mkdtemp()
    ckpt = Path(tmp) / "vgg19-random.pth"
    params = save_checkpoint(ckpt)
    print("--- by extent (all modules) ---")
    for h, w in [(90, 122), (90, 128), (96, 122), (88, 122), (90, 120), (92, 124), (45, 61), (94, 126), (86, 118), (100, 100),
                 (724, 724)]:
        e, at = run(h, w, tmp, params, ckpt)
        print(f"{h}x{w}: rel {e:.2e} at {at}")
    print("--- 90x122 by module ---")
    for name, over in [
        ("tv only", dict(content_weight=0.0, style_weight=0.0)),
        ("content only", dict(style_weight=0.0, tv_weight=0.0)),
        ("style relu1_1", dict(content_weight=0.0, tv_weight=0.0, style_layers="relu1_1")),
        ("style relu2_1", dict(content_weight=0.0, tv_weight=0.0, style_layers="relu2_1")),
        ("style relu3_1", dict(content_weight=0.0, tv_weight=0.0, style_layers="relu3_1")),
        ("style relu4_1", dict(content_weight=0.0, tv_weight=0.0, style_layers="relu4_1")),
        ("style relu5_1", dict(content_weight=0.0, tv_weight=0.0, style_layers="relu5_1")),
        ("style relu1_2 (pre-pool)", dict(content_weight=0.0, tv_weight=0.0, style_layers="relu1_2")),
        ("style relu2_2 (pre-pool)", dict(content_weight=0.0, tv_weight=0.0, style_layers="relu2_2")),
        ("avg pooling, all", dict(pooling="avg")),
    ]:
        try:
            e, at = run(90, 122, tmp, params, ckpt, **over)
            print(f"{name}: rel {e:.2e} at {at}")
        except Exception as ex:  # noqa: BLE001
            print(f"{name}: {type(ex).__name__}: {ex}")


if __name__ == "__main__":
    main()
