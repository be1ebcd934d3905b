"""Tile-shape sweep of conv_tc_kernel on the layer shapes of the small scales (256^2 / 512^2 VGG-19: where tiles < SMs and
the launch is bound by L2 -> shared-memory operand traffic and pipeline latency rather than by the tensor core), in ONE
process: MAUA_CONV_FORCE is read at every launch.  Prints, per shape, the time of every (BN, MT, CG) and the shape the
built-in chooser picks.   python tools/sweep_conv_small.py [reps]"""
import os
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from maua_style_b200 import _lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 15
lib = _lib.load()
_lib.require_gpu()
SHAPES = [(512, 512, 16, 16), (512, 512, 32, 32), (256, 512, 32, 32), (512, 512, 64, 64), (256, 512, 64, 64),
          (256, 256, 64, 64), (128, 256, 64, 64), (256, 256, 128, 128), (128, 256, 128, 128), (128, 128, 128, 128),
          (64, 128, 128, 128), (128, 128, 256, 256), (64, 64, 256, 256), (512, 512, 128, 128)]
if os.environ.get("SWEEP_BIG"):  # the layer shapes of the 1024^2 scale (full waves: what the chooser's shape_rate table is fitted to)
    SHAPES = [(64, 64, 1024, 1024), (64, 128, 512, 512), (128, 128, 512, 512), (128, 256, 256, 256), (256, 256, 256, 256),
              (256, 512, 128, 128), (512, 512, 128, 128), (512, 512, 64, 64)]
CFGS = [(bn, mt, cg) for bn in (256, 128, 64, 32) for mt in (2, 1) for cg in (2, 1)]
flush = torch.empty(64 << 20, device="cuda")  # 256 MB: evict L2 between timed launches


def time_cfg(x, wg, b, y, cin, cout, h, w, force):
    if force:
        os.environ["MAUA_CONV_FORCE"] = ",".join(str(v) for v in force)
    else:
        os.environ.pop("MAUA_CONV_FORCE", None)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.maua_conv3x3_fwd(_lib.ptr(x), _lib.ptr(wg), _lib.ptr(b), _lib.ptr(y), 1, h, w, cin, cout, 1, 0, _lib.stream_ptr()))
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


for cin, cout, h, w in SHAPES:
    x = torch.randn(1, h, w, cin, device="cuda")
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    b = torch.randn(cout, device="cuda")
    wg = torch.empty(cout, 9 * cin, device="cuda")
    _lib.check(lib.maua_prep_conv_weights(_lib.ptr(wt), _lib.ptr(wg), cout, cin, 0, _lib.stream_ptr()))
    y = torch.empty(1, h, w, cout, device="cuda")
    base = time_cfg(x, wg, b, y, cin, cout, h, w, None)
    res = []
    for cfg in CFGS:
        if cout % cfg[0]:
            continue
        res.append((time_cfg(x, wg, b, y, cin, cout, h, w, cfg), cfg))
    res.sort()
    fl = 2.0 * 9 * cin * cout * h * w
    print(f"conv {cin}->{cout} {h}x{w} ({fl / 1e9:.2f} GFLOP): chooser {base:.1f} us | " +
          "  ".join(f"{','.join(str(v) for v in c)}: {t:.1f}" for t, c in res[:int(os.environ.get("SWEEP_TOP", "6"))]), flush=True)
os.environ.pop("MAUA_CONV_FORCE", None)
