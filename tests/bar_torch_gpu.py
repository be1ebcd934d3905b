#!/usr/bin/env python
"""The "bar": the reference's algorithm executed by stock PyTorch (cuDNN conv, cuBLAS fp32 mm, ATen, torch.optim) on
the same B200 -- what `style.py --gpu 0` would run today.  NOT a test and NOT part of bench.py's contract: a
measurement script (it drives the CPU oracle's torch restatement of the reference on a CUDA device, with
torch.optim.Adam / LBFGS exactly as optim.py:180-196 configures them, including the per-module host syncs of
optim.py:210).

    python tests/bar_torch_gpu.py [--size 1024] [--iters 30] [--optimizer lbfgs|adam]   ->  one JSON line
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import maua_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--optimizer", default="lbfgs")
    ap.add_argument("--tf32-matmul", action="store_true", help="allow TF32 in torch.mm too (default: fp32, as torch ships)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True  # style.py:19
    torch.backends.cuda.matmul.allow_tf32 = a.tf32_matmul
    params = [(w.to(dev), b.to(dev)) for w, b in O.he_init_vgg19(0)]
    cfg = O.StyleConfig(content_weight=5.0, optimizer=a.optimizer)
    net = O.OracleNet(params, cfg)
    S = a.size
    O.set_content_targets(net, O.synthetic_image(S, S, seed=1, smooth=True).to(dev))
    O.set_style_targets(net, [O.synthetic_image(S, S, seed=2).to(dev)], [1.0])
    for m in net.losses:
        m.mode = "loss"
    pastiche = torch.nn.Parameter((O.synthetic_image(S, S, seed=4) * 0.25).to(dev))
    count = [0]

    def closure():  # optim.py:201-221
        pastiche.grad = None
        net(pastiche)
        total = 0
        for m in net.losses:
            if isinstance(m.loss, int):
                continue
            m.loss.detach().cpu().item()  # optim.py:210 (log_losses bookkeeping: one host sync per module)
            total = total + m.loss
        total.backward()
        for m in net.losses:
            m.loss = 0
        count[0] += 1
        return total

    def run(n):
        if a.optimizer == "lbfgs":
            opt = torch.optim.LBFGS([pastiche], max_iter=n, tolerance_change=-1, tolerance_grad=-1, history_size=100)
            opt.step(closure)
        else:
            opt = torch.optim.Adam([pastiche], lr=1.0)
            for _ in range(n):
                opt.step(closure)

    run(5)  # warm-up (cuDNN autotune)
    torch.cuda.synchronize()
    count[0] = 0
    t0 = time.perf_counter()
    run(a.iters)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"bar": "stock PyTorch (cuDNN/cuBLAS/ATen + torch.optim) running the reference algorithm on this GPU",
                      "size": S, "optimizer": a.optimizer, "iters": count[0], "it_per_s": count[0] / dt,
                      "ms_per_iter": dt / count[0] * 1e3, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
                      "tf32_conv": torch.backends.cudnn.allow_tf32, "tf32_matmul": a.tf32_matmul,
                      "note": "L-BFGS history ramps from 0 within the timed step() (reference behaviour)"}))


if __name__ == "__main__":
    main()
