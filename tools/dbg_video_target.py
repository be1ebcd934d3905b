"""Debug: per-block error of the exact-mode dynamic video target against the CPU oracle (img_vid golden inputs)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from helpers import O, load_golden, make_args, save_checkpoint, video_cfg, video_inputs  # noqa: E402

from maua_style_b200 import _lib, models, optim  # noqa: E402

name = "img_vid_windows_adam_48x64"
z, meta = load_golden(name)
tmp = Path("/tmp/dbgv"); tmp.mkdir(exist_ok=True)
path = tmp / "vgg19-random.pth"
params = save_checkpoint(path)
args = make_args(path, tmp, transfer_type="img_vid", gram_frame_window=meta["gfw"], avg_frame_window=meta["afw"], **meta["over"])
net, losses = models.load_model(args)
net.set_impl(_lib.MAUA_IMPL_FP32)
content, styles, init = video_inputs(meta)
optim.set_content_targets(net, content, args)
optim.set_style_video_targets(net, [s.cuda() for s in styles], args)
cfg = video_cfg(meta)
onet = O.OracleNet(params, cfg)
O.set_content_targets(onet, content)
O.set_style_video_targets(onet, styles, cfg.blend(len(styles)), meta["gfw"])
B = meta["gfw"]
for i, (m, om) in enumerate(zip(net.style_losses, onet.style_losses)):
    vt, ovt = m.video_target.cpu().double(), om.video_target.double()
    C_ = vt.shape[0] // B
    print(f"style {i}: C {C_} total rel {float((vt - ovt).norm() / ovt.norm()):.2e}  static rel {float((m.target.cpu().double() - om.target.double()).norm() / om.target.double().norm()):.2e}")
    for a in range(B):
        print("   ", " ".join(f"{float((vt[a*C_:(a+1)*C_, b*C_:(b+1)*C_] - ovt[a*C_:(a+1)*C_, b*C_:(b+1)*C_]).norm() / ovt[a*C_:(a+1)*C_, b*C_:(b+1)*C_].norm()):.2e}" for b in range(B)))
