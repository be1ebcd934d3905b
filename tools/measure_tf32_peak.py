"""cuBLAS TF32 / bf16 GEMM throughput on this GPU (context for the conv roofline denominator):  one JSON line."""
import json, torch
torch.backends.cuda.matmul.allow_tf32 = True
N = 8192
out = {}
for name, dt in (("tf32", torch.float32), ("bf16", torch.bfloat16)):
    a = torch.randn(N, N, device="cuda", dtype=dt); b = torch.randn(N, N, device="cuda", dtype=dt)
    for _ in range(3): a @ b
    torch.cuda.synchronize()
    best = 0.0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    for i in range(10):
        ev[i].record(); a @ b
    ev[10].record(); torch.cuda.synchronize()
    for i in range(10):
        best = max(best, 2 * N**3 / (ev[i].elapsed_time(ev[i + 1]) * 1e-3) / 1e12)
    # sustained: back to back for ~2 s
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 400 if dt == torch.float32 else 800
    e0.record()
    for _ in range(reps): a @ b
    e1.record(); torch.cuda.synchronize()
    out[name] = {"burst_tflops": round(best, 1), "sustained_tflops": round(reps * 2 * N**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12, 1)}
print(json.dumps({"cublas_gemm_8192": out, "gpu": torch.cuda.get_device_name(0)}))
