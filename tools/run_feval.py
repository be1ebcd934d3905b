"""A few fevals of one configuration for ncu captures: python tools/run_feval.py {nin|prune|vgg19|window} SIZE"""
import sys
import tempfile
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from maua_style_b200 import models, optim, synthetic as O  # noqa: E402

kind, size = sys.argv[1], int(sys.argv[2])
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
tmp = tempfile.mkdtemp(prefix="maua_feval_")
over = dict(optimizer="adam", gpu="0")
if kind == "nin":
    ckpt = Path(tmp) / "nin-random.pth"
    O.save_random_checkpoint(ckpt, channels=models.NIN_LAYERS)
    over.update(style_layers="relu1,relu3,relu5,relu7,relu9,relu11", content_layers="relu8")
elif kind == "prune":
    ckpt = Path(tmp) / "vgg16-prune-random.pth"
    O.save_random_checkpoint(ckpt, channels=models.channel_list["VGG-16p"])
else:
    ckpt = Path(tmp) / "vgg19-random.pth"
    O.save_random_checkpoint(ckpt)
frames = 4 if kind == "window" else 1
if kind == "window":
    over.update(transfer_type="img_vid", gram_frame_window=frames, avg_frame_window=-1)
a = O.reference_args(ckpt, tmp, **over)
net, losses = models.load_model(a)
optim.set_content_targets(net, O.synthetic_image(size, size, seed=1, smooth=True).to(dev), a)
if kind == "window":
    video = torch.cat([O.synthetic_image(size, size, seed=20 + f) for f in range(frames + 1)]).to(dev)
    optim.set_style_video_targets(net, [video], a)
else:
    optim.set_style_targets(net, [O.synthetic_image(size, size, seed=2).to(dev)], a)
for m in losses:
    m.mode = "loss"
x = torch.cat([O.synthetic_image(size, size, seed=40 + f) * 0.25 for f in range(frames)]).to(dev).contiguous()
up = torch.zeros(net._n_slots, device=dev)
up[net._live_slots()] = 1.0
for _ in range(3):
    net._forward_plan(x, keep=True)
    g = net._backward_plan(up)
torch.cuda.synchronize()
print("feval ok", kind, size, float(net._loss_vec.sum()), float(g.abs().mean()))
