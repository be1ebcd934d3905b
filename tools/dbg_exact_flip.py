"""Find the pooling windows / ReLU signs whose decision differs between the exact plan and the fp64 oracle (90x122 golden
inputs) and print the values involved.  Usage (GPU box): python tools/dbg_exact_flip.py"""
import sys
import tempfile
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import O, make_args, save_checkpoint  # noqa: E402

from maua_style_b200 import _lib, models, optim  # noqa: E402

h, w = 90, 122
tmp = tempfile.mkdtemp()
ckpt = Path(tmp) / "vgg19-random.pth"
params = save_checkpoint(ckpt)
layers = "relu1_2,relu2_2,relu3_4,relu4_4"
args = make_args(ckpt, tmp, temporal_weight=0.0, style_layers=layers, content_layers="relu5_1")
net, losses = models.load_model(args)
net.set_impl(_lib.MAUA_IMPL_FP32)
init = (O.synthetic_image(h, w, seed=4) * 0.25)
style = O.synthetic_image(h, w, seed=2)
optim.set_content_targets(net, init, args)
optim.set_style_targets(net, [style], args)
for m in losses:
    m.mode = "loss"
optim.feval(net, init.clone().cuda())
cfg = O.StyleConfig(temporal_weight=0.0, style_layers=layers, content_layers="relu5_1")
for dt, tag in ((torch.float64, "fp64"), (torch.float32, "fp32 CPU")):
    onet = O.OracleNet([(a.to(dt), b.to(dt)) for a, b in params], cfg)
    taps = {}
    onet(init.to(dt), taps=taps)
    names = O.relu_names(O.VGG19_CHANNELS)
    for t, (ridx, mod) in enumerate(net.taps):
        nm = names[ridx]
        if nm == "relu5_1":
            continue
        ours = net.tap_feature(t).cpu()
        ref = taps[nm]
        _, io = F.max_pool2d(ours, 2, 2, return_indices=True)
        _, ir = F.max_pool2d(ref, 2, 2, return_indices=True)
        diff = (io != ir)
        sign = ((ours > 0) != (ref > 0))
        print(f"[{tag}] {nm}: {int(diff.sum())} of {diff.numel()} windows pick another arg-max, {int(sign.sum())} ReLU signs differ, "
              f"feature rel err {float((ours.double() - ref.double()).norm() / ref.double().norm()):.2e}")
        for idx in diff.nonzero()[:5]:
            b, c, ph, pw = [int(v) for v in idx]
            wo = ours[b, c, 2 * ph:2 * ph + 2, 2 * pw:2 * pw + 2].flatten().tolist()
            wr = ref[b, c, 2 * ph:2 * ph + 2, 2 * pw:2 * pw + 2].flatten().tolist()
            print(f"    channel {c} window ({ph},{pw}) ours {['%.9g' % v for v in wo]}  {tag} {['%.12g' % v for v in wr]}")
        for idx in sign.nonzero()[:5]:
            b, c, y, x = [int(v) for v in idx]
            print(f"    channel {c} pixel ({y},{x}) ours {float(ours[b, c, y, x]):.9g}  {tag} {float(ref[b, c, y, x]):.12g}")
