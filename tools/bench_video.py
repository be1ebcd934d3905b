"""A measured job through the vid_img driver (style.vid_img_tensors; style.py:145-300): N consecutive synthetic frames at SIZE^2,
one scale, PASSES passes of ITERS // PASSES L-BFGS iterations per frame, flow fields and reliability maps as inputs.  Wall clock
around the whole job (targets, warps, temporal captures, optimisation, 8-bit results), device synchronised on both sides; the job
runs twice and the second (warm plan / graph caches) run is reported next to the first.

    python tools/bench_video.py [--frames 8] [--size 1024] [--iters 100] [--passes 2]
"""
import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from maua_style_b200 import _lib, style, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--optimizer", default="lbfgs")
    o = ap.parse_args()
    _lib.require_gpu()
    tmp = Path(tempfile.mkdtemp(prefix="maua_bench_video_"))
    ckpt = tmp / "vgg19-random.pth"
    synthetic.save_random_checkpoint(ckpt)
    S = o.size
    frames = [synthetic.synthetic_image(S, S, seed=100 + i, smooth=True).cuda() for i in range(o.frames)]
    styles = [synthetic.synthetic_image(S, S, seed=2).cuda()]
    g = torch.Generator().manual_seed(5)
    flow = (torch.randn(S // 4, S // 4, 2, generator=g) * 0.002)       # normalised + blurred field (style.read_flo's output)
    rel = (torch.rand(1, 1, S // 4, S // 4, generator=g) > 0.1).float().cuda()
    flows = lambda d, i, j: (flow if d == "forward" else -flow, rel)

    def job():
        a = synthetic.reference_args(ckpt, tmp, transfer_type="vid_img", optimizer=o.optimizer, image_sizes=[S], num_iters=[o.iters],
                                     passes_per_scale=o.passes, init="content", temporal_blend=0.5, loop=False, style_scale=1.0,
                                     match_histograms=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        store = style.vid_img_tensors(frames, styles, a, flows)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, store

    cold, _ = job()
    warm, store = job()
    evals = o.frames * o.passes * (o.iters // o.passes)
    print(json.dumps({"workload": f"vid_img driver: {o.frames} frames of {S}x{S}, 1 scale, {o.passes} passes x {o.iters // o.passes} {o.optimizer} "
                                  "iterations per frame, temporal loss on from pass 2, synthetic frames / flows",
                      "seconds_cold": cold, "seconds": warm, "frames_per_min": 60.0 * o.frames / warm, "iterations": evals,
                      "value": evals / warm, "unit": "it/s", "results": len(store),
                      "timing": "wall clock around style.vid_img_tensors, device synchronised on both sides"}), flush=True)


if __name__ == "__main__":
    main()
