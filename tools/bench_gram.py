"""Gram SYRK + finalize (csrc/gram.cu, maua_gram) alone on the five style-tap shapes of a VGG-19 at S x S: us per call,
back-to-back launches (host overhead hidden behind the queue).   python tools/bench_gram.py [S...]"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from maua_style_b200 import _lib

lib = _lib.load()
_lib.require_gpu()
for S in [int(a) for a in sys.argv[1:]] or [1024, 256]:
    out = []
    for c, div in ((64, 1), (128, 2), (256, 4), (512, 8), (512, 16)):
        hw = (S // div) ** 2
        f = torch.randn(hw, c, device="cuda")
        G = torch.empty(c, c, device="cuda")
        mean = torch.empty(c, device="cuda")
        ws = torch.empty(int(lib.maua_gram_workspace_bytes(c)), dtype=torch.uint8, device="cuda")
        for timed in (0, 1):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30):
                _lib.check(lib.maua_gram(_lib.ptr(f), C.c_long(hw), c, 0, _lib.ptr(G), _lib.ptr(mean), _lib.ptr(ws), 0, _lib.stream_ptr()))
            e1.record()
            e1.synchronize()
        out.append(f"C{c}/P{hw}: {e0.elapsed_time(e1) * 1e3 / 30:.1f}")
    print(f"gram {S}^2: " + "  ".join(out), flush=True)
