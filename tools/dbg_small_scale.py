"""Where do the first iterations of an optimize() call at 256^2 go?  Per-10-iteration wall times (device synchronised) of
GraphedIteration through optimize_device's own loop state, 3 calls in a row."""
import sys
import tempfile
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from maua_style_b200 import models, optim, synthetic as O  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
tmp = tempfile.mkdtemp(prefix="maua_dbg_")
ckpt = Path(tmp) / "vgg19-random.pth"
O.save_random_checkpoint(ckpt)
a = O.reference_args(ckpt, tmp, optimizer="lbfgs", gpu="0")
content = O.synthetic_image(size, size, seed=1, smooth=True).to(dev)
sty = O.synthetic_image(size, size, seed=2, smooth=True).to(dev)
for call in range(3):
    net, losses = models.load_model(a)
    net.reuse_target_buffers = True
    optim.set_content_targets(net, content, a)
    optim.set_style_targets(net, [sty], a)
    for m in losses:
        m.mode = "loss"
    live = net._live_slots()
    state = optim._loop_state(net, content.clone(), a, live)
    it = state.iteration
    torch.cuda.synchronize()
    marks = []
    t0 = time.perf_counter()
    for i in range(1, 301):
        it()
        if i <= 10 or i % 10 == 0:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            marks.append((i, 1e3 * (t1 - t0)))
            t0 = t1
    print(f"call {call} (graph {'kept' if it.graph is not None else 'none'}): " + " ".join(f"{i}:{ms:.2f}" for i, ms in marks[:10]))
    print("   per 10: " + " ".join(f"{i}:{ms:.1f}" for i, ms in marks[10:]))
    it.net = None
