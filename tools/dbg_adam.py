import ctypes as C, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from maua_style_b200 import _lib
lib = _lib.load(); _lib.require_gpu()
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
for trial in range(3):
    g = torch.Generator().manual_seed(7 + trial)
    n = 3 * 37 * 41 + 3
    p = torch.randn(n, generator=g).requires_grad_(True)
    opt = torch.optim.Adam([p], lr=1.0)
    pd = p.detach().clone().cuda(); pd2 = pd.clone()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    m2, v2 = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    sd = torch.zeros(1, dtype=torch.int32, device="cuda")
    for step in range(1, 6):
        grad = torch.randn(n, generator=g) * 10 ** (step - 3)
        p.grad = grad.clone(); opt.step()
        gradd = grad.cuda()
        _lib.check(lib.maua_adam_step(_lib.ptr(pd), _lib.ptr(gradd), _lib.ptr(m), _lib.ptr(v), C.c_long(n), C.c_float(1.0), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8), step, _lib.stream_ptr()))
        sd.add_(1)
        _lib.check(lib.maua_adam_step_dev(_lib.ptr(pd2), _lib.ptr(gradd), _lib.ptr(m2), _lib.ptr(v2), C.c_long(n), C.c_float(1.0), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8), _lib.ptr(sd), _lib.stream_ptr()))
        torch.cuda.synchronize()
        print(trial, step, "host-step rel", rel(pd, p.detach()), "dev-step rel", rel(pd2, p.detach()), "max abs", float((pd.cpu()-p.detach()).abs().max()))
