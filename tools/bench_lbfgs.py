"""L-BFGS step alone (csrc/lbfgs.cu) at full history: ms per step and effective HBM rate for n = 3 * S * S.
   python tools/bench_lbfgs.py [sizes...]"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from maua_style_b200 import _lib

lib = _lib.load()
_lib.require_gpu()
sizes = [int(a) for a in sys.argv[1:]] or [256, 512, 1024]
K = 100
for S in sizes:
    n = 3 * S * S
    g = torch.Generator(device="cuda").manual_seed(S)
    scale = torch.rand(n, device="cuda", generator=g) * 0.6 + 0.7   # gradient of a convex quadratic: every pair is accepted
    x = torch.randn(n, device="cuda", generator=g)
    state = C.c_void_p()
    _lib.check(lib.maua_lbfgs_create(C.c_long(n), K, C.c_float(1.0), C.c_float(-1.0), C.byref(state)))
    ts = []
    for it in range(K + 40):
        grad = x * scale + 1e-3 * torch.randn(n, device="cuda", generator=g)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.maua_lbfgs_step(state, _lib.ptr(x), _lib.ptr(grad), _lib.stream_ptr()))
        e1.record()
        e1.synchronize()
        if it >= K + 10:
            ts.append(e0.elapsed_time(e1))
    n_iter, hist, halted = C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.maua_lbfgs_query(state, C.byref(n_iter), C.byref(hist), C.byref(halted), _lib.stream_ptr()))
    lib.maua_lbfgs_destroy(state)
    ts.sort()
    ms = ts[len(ts) // 2]
    gb = 16.0 * hist.value * n / 1e9
    print(f"lbfgs step {S}^2: history {hist.value} halted {halted.value}: {ms * 1e3:.1f} us, {gb / ms:.0f} GB/s of history traffic", flush=True)
