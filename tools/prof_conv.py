"""Run one conv3x3 layer through the C ABI a few times (for ncu):  python tools/prof_conv.py CIN COUT H W [reps]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from maua_style_b200 import _lib

cin, cout, h, w = map(int, sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
lib = _lib.load()
_lib.require_gpu()
x = torch.randn(1, h, w, cin, device="cuda")
wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
b = torch.randn(cout, device="cuda")
wg = torch.empty(cout, 9 * cin, device="cuda")
_lib.check(lib.maua_prep_conv_weights(_lib.ptr(wt), _lib.ptr(wg), cout, cin, 0, _lib.stream_ptr()))
y = torch.empty(1, h, w, cout, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
for i in range(reps):
    ev[i].record()
    _lib.check(lib.maua_conv3x3_fwd(_lib.ptr(x), _lib.ptr(wg), _lib.ptr(b), _lib.ptr(y), 1, h, w, cin, cout, 1, 0, _lib.stream_ptr()))
ev[reps].record()
torch.cuda.synchronize()
fl = 2.0 * 9 * cin * cout * h * w
for i in range(reps):
    ms = ev[i].elapsed_time(ev[i + 1])
    print(f"conv {cin}->{cout} {h}x{w}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s")
