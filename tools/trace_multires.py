"""The multi-resolution job of bench.py (256 -> 512 -> 1024, L-BFGS 500 / 400 / 200) run twice with MAUA_TRACE=1: the second,
warm run shows where the wall-clock time of the whole job goes besides the iterations (set-up per scale, target capture,
loop state, graph capture)."""
import sys
import tempfile
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from maua_style_b200 import image_ops, style, synthetic as O  # noqa: E402

import gc

_gc_t = {}


def _gc_cb(phase, info):  # how long does the cyclic collector stop the host thread, and when?
    if phase == "start":
        _gc_t["t"] = time.perf_counter()
    else:
        dt = 1e3 * (time.perf_counter() - _gc_t.get("t", time.perf_counter()))
        if dt > 2.0:
            print(f"[trace] gc generation {info['generation']}: {dt:.1f} ms, {info['collected']} collected", file=sys.stderr, flush=True)


gc.callbacks.append(_gc_cb)
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
tmp = tempfile.mkdtemp(prefix="maua_trace_")
ckpt = Path(tmp) / "vgg19-random.pth"
O.save_random_checkpoint(ckpt)
a = O.reference_args(ckpt, tmp, optimizer="lbfgs", gpu="0")
a.image_sizes, a.num_iters, a.init, a.style_scale = [256, 512, 1024], [500, 400, 200], "content", 1.0
content = O.synthetic_image(1024, 1024, seed=1, smooth=True).to(dev)
sty = O.synthetic_image(1024, 1024, seed=2, smooth=True).to(dev)
for run in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = style.img_img_tensors(content, [sty], a)
    torch.cuda.synchronize()
    print(f"[trace] run {run}: whole job {time.perf_counter() - t0:.3f} s", file=sys.stderr, flush=True)
