#!/usr/bin/env python
"""bench.py -- VGG-19 style-transfer iterations/sec on B200 (the BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--optimizer lbfgs|adam] [--impl ours|reference]

Workload (BASELINE.json configs[1], final scale): one 1024x1024 image, VGG-19 to relu5_1, Gram StyleLoss on
relu1_1..relu5_1, ContentLoss on relu4_2, TVLoss, L-BFGS (history 100, no line search) pixel update; synthetic
inputs U(0,255)-BGR mean, He-normal random-init weights (no checkpoints offline).  One "step" = one feval
(forward + all losses + backward to the image) + one optimizer update = one reference iteration (optim.py:201-241).
The L-BFGS history is filled to 100 pairs during set-up so that the timed steps run at full history.

  value    whole-job iterations/sec with the pastiche resident in HBM (N GPUs: N independent images, one per rank,
           no collective -- BASELINE.json config 5 sharding)
  e2e      the same step driven with HOST buffers: every step copies the pastiche host->device from pinned memory
           and reads the updated pastiche + total loss back device->host
  roofline conv3x3 implicit-GEMM kernels (forward + dgrad, the dominant kernel family): algorithmic FLOPs
            / CUDA-event time of those launches, against the measured tensor peak
  cpu_baseline  the CPU oracle (port of the reference algorithm) on the host cores, bounded sample

`--impl reference` times the reference's CPU implementation of the same step (the oracle port: the reference is a
Python package that cannot travel to the GPU box) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

METRIC = "vgg19_style_iterations_per_sec"
UNIT = "it/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--optimizer", default="lbfgs", choices=["lbfgs", "adam"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--history-prefill", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-launch profile JSON here")
    ap.add_argument("--no-multires", action="store_true", help="skip the 256->512->1024 multi-resolution job (configs[1] whole)")
    ap.add_argument("--streams", type=int, default=1,
                    help="experimental: S independent images per GPU on S streams (batch jobs, BASELINE.json configs[4]); "
                         "prints its own JSON line, the default line is unchanged")
    return ap.parse_args()


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    bf16_burst=float(d["bf16_tflops"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1400.0, bf16_burst=1590.0, source="fallback (B200_PROFILING.md)")


def conv_traffic_per_launch(size):
    """Average DRAM bytes per conv_tc launch from the committed `ncu --set full` capture of the same workload
    (tools/ncu_table.py --traffic writes profiles/conv_traffic.json); None when no capture exists for this size."""
    f = ROOT / "profiles" / "conv_traffic.json"
    try:
        d = json.loads(f.read_text())
        return d.get(str(size), {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def workload_config(size, optimizer):
    return {
        "workload": f"VGG-19 Gram style transfer {size}x{size}, content relu4_2 + style relu1_1..relu5_1 + TV, "
                    f"{optimizer} (BASELINE.json configs[1] final scale)",
        "image": [size, size], "styles": 1, "optimizer": optimizer, "lbfgs_history": 100,
        "weights": "He-normal random init (seed 0)", "timing": "inputs larger than L2: 1.2 GB of activations per step",
        "sharding": "one independent image per GPU, no collective",
    }


# ---------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_step_fn(size, optimizer):
    from oracle import maua_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_flush_denormal(True)
    params = O.he_init_vgg19(0)
    cfg = O.StyleConfig(content_weight=5.0, optimizer=optimizer)
    net = O.OracleNet(params, cfg)
    content = O.synthetic_image(size, size, seed=1, smooth=True)
    style = O.synthetic_image(size, size, seed=2)
    O.set_content_targets(net, content)
    O.set_style_targets(net, [style], [1.0])
    for m in net.losses:
        m.mode = "loss"
    state = {"p": O.synthetic_image(size, size, seed=4) * 0.25, "m": None, "v": None, "t": 0}

    def step():  # feval + a plain Adam-style pixel update (the optimizer is <1% of a CPU iteration)
        _, _, g = O.feval(net, state["p"])
        state["t"] += 1
        if state["m"] is None:
            state["m"], state["v"] = torch.zeros_like(g), torch.zeros_like(g)
        state["m"].lerp_(g, 0.1)
        state["v"].mul_(0.999).addcmul_(g, g, value=0.001)
        state["p"] = state["p"] - state["m"] / (state["v"].sqrt() + 1e-8)

    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size = args.size
    step = cpu_step_fn(size, args.optimizer)
    t0 = time.perf_counter()
    step()
    probe = time.perf_counter() - t0
    scale, sample = 1.0, f"full {size}x{size} iterations"
    budget = 150.0
    if probe * (args.steps + args.warmup) > budget and size > 256:
        # bounded sample: same step on a centre crop with 1/4 (or 1/16) of the pixels; CPU conv cost is linear in
        # pixels, so throughput is scaled back by the pixel ratio
        sub = size // 2 if probe * (args.steps + args.warmup) / 4 <= budget else size // 4
        scale = (sub * sub) / float(size * size)
        step = cpu_step_fn(sub, args.optimizer)
        sample = f"{sub}x{sub} iterations scaled by {scale:.4f} to {size}x{size}-equivalent (probe {probe:.1f} s/iter at full size)"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps / dt * scale
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3 / scale, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(size, args.optimizer),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(size, optimizer):
    """Bounded sample (about 10-30 s of CPU work) of the same step on the host cores."""
    sub = size
    step = cpu_step_fn(sub, optimizer)
    t0 = time.perf_counter()
    step()
    probe = time.perf_counter() - t0
    scale = 1.0
    sample = f"2 iterations at {size}x{size} after 1 warm-up"
    if probe > 12.0 and size >= 512:
        sub = size // 2
        scale = 0.25
        step = cpu_step_fn(sub, optimizer)
        step()
        sample = f"2 iterations at {sub}x{sub} scaled by 0.25 to {size}x{size}-equivalent (full-size probe {probe:.1f} s)"
    t0 = time.perf_counter()
    for _ in range(2):
        step()
    dt = time.perf_counter() - t0
    return {"value": 2 / dt * scale, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.proc = None
        self.gpu = gpu_index
        self.path = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock and active throttle reasons of the samples taken inside [t_begin, t_end] (wall clock)."""
        import datetime

        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for ln in open(self.path).read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]), parts[4:8]))
            except ValueError:
                continue
        inside = [r for r in rows if t_begin is None or (t_begin - 0.05 <= r[0] <= t_end + 0.05)]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: use every sample taken under load
            inside, window = rows, "whole run (timed region shorter than the sampling period)"
        sm, mx, reasons = [], [], set()
        for _, a, b, flags in inside:
            sm.append(a); mx.append(b)
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], flags):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def lbfgs_launches(calls, hist):
    return 4  # multi-dot pass, partial reduce, scalar recurrences, direction + update (maua_style_b200/csrc/lbfgs.cu)


def run_ours(args):
    import torch.distributed as dist

    from maua_style_b200 import _lib, models, optim, synthetic as O  # the product path never touches oracle/

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.require_gpu()

    size, K, W = args.size, args.steps, args.warmup
    tmp = tempfile.mkdtemp(prefix=f"maua_bench_{rank}_")
    ckpt = Path(tmp) / "vgg19-random.pth"
    O.save_random_checkpoint(ckpt)
    a = O.reference_args(ckpt, tmp, optimizer=args.optimizer, gpu=str(local_rank))
    net, losses = models.load_model(a)
    content = O.synthetic_image(size, size, seed=1 + 10 * rank, smooth=True)
    style = O.synthetic_image(size, size, seed=2)
    init = O.synthetic_image(size, size, seed=4 + 10 * rank) * 0.25
    optim.set_content_targets(net, content.to(dev), a)
    optim.set_style_targets(net, [style.to(dev)], a)
    for m in losses:
        m.mode = "loss"
    pastiche = init.to(dev).contiguous()
    hist = 100
    opt = optim.PixelOptimizer(pastiche, args.optimizer, lr=1.0, history=hist)
    up = torch.zeros(net._n_slots, device=dev)
    live = net._live_slots()
    up[live] = 1.0
    # the iteration exactly as optim.optimize drives it: 2 eager rounds, then one CUDA-graph replay per iteration
    step = optim.GraphedIteration(net, pastiche, opt, up)

    sampler = ClockSampler(local_rank)
    sampler.start()
    # set-up: fill the L-BFGS history so the timed steps run at full history (not part of warm-up or timing)
    if args.optimizer == "lbfgs":
        for _ in range(args.history_prefill):
            step()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    for _ in range(W):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    calls0 = opt.step_count
    t_begin = time.time()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * K / (ms_total / 1e3)
    fwd_l, bwd_l = net.last_launches()
    opt_l = K * (lbfgs_launches(calls0, hist) if args.optimizer == "lbfgs" else 2)  # adam: step counter + update
    gpu_launches = K * (fwd_l + bwd_l) + opt_l  # kernels executed (as nodes of the replayed graph when graphs are on)

    # ---- end to end with host buffers (pinned), every step H2D pastiche + D2H result ----
    host = [pastiche.detach().cpu().pin_memory(), torch.empty_like(pastiche, device="cpu").pin_memory()]
    host_loss = torch.empty(1).pin_memory()
    total_dev = torch.zeros(1, device=dev)

    def step_e2e():
        host_in, host_out = host
        pastiche.copy_(host_in, non_blocking=True)       # H2D: this step's input image from pinned host memory
        step()
        torch.sum(net._loss_vec, dim=0, keepdim=True, out=total_dev)
        host_out.copy_(pastiche, non_blocking=True)      # D2H: the updated image ...
        host_loss.copy_(total_dev, non_blocking=True)    # ... and the total loss
        torch.cuda.current_stream().synchronize()        # the caller holds the result on the host
        host[0], host[1] = host_out, host_in             # next step's input is this step's result (host side)

    for _ in range(min(W, 3)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * K / float(dt.item())
    nbytes = pastiche.numel() * 4

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32", "dtype_detail": "TF32 tensor-core operands, fp32 accumulate, fp32 storage", "data": "synthetic",
        "config": workload_config(size, args.optimizer), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes + 4},
        "gpu_launches": gpu_launches,
        "images_per_min_at_1000_iters": value * 60.0 / 1000.0,
        "cuda_graph": bool(step.graph is not None),
    }

    if rank == 0:
        # ---- roofline of the dominant kernel family, measured live with per-launch CUDA events ----
        pk = peaks()
        net.set_profile(True)
        recs = []
        for _ in range(3):
            net._forward_plan(pastiche, keep=True)
            net._backward_plan(up)
            recs.append(net.profile())
        net.set_profile(False)
        prof = recs[-1]
        conv = [r for r in prof if r["name"] in ("conv_fwd", "conv_dgrad")]
        conv_ms = sum(r["ms"] for r in conv)
        conv_flops = sum(r["flops"] for r in conv)
        iter_ms = sum(r["ms"] for r in prof)
        tf32_peak = pk["bf16"] / 2.0
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        line["roofline"] = {
            "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv3x3, forward + dgrad, 25 launches/iteration)",
            "bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
            "peak_source": f"{pk['source']}: bf16_tflops_sustained {pk['bf16']:.1f} / 2 -- TF32 MMA issues at half the bf16 rate",
            "frac_of_bf16_peak": achieved / pk["bf16"],
            "traffic": conv_traffic_per_launch(size),
            "traffic_source": "profiles/conv_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per conv_tc launch, ncu --set full)",
            "achieved_per_launch_flops": conv_flops / max(len(conv), 1), "launches_per_iteration": len(conv),
            "share_of_feval": conv_ms / iter_ms if iter_ms > 0 else None,
            "algorithmic_flops_per_iteration": conv_flops,
        }
        by = {}
        for r in prof:
            d = by.setdefault(r["name"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
            d["ms"] += r["ms"]; d["flops"] += r["flops"]; d["bytes"] += r["bytes"]; d["n"] += 1
        line["kernel_breakdown_ms"] = {k: round(v["ms"], 4) for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"])}
        mem = [(k, v) for k, v in by.items() if v["flops"] == 0 and v["bytes"] > 0 and v["ms"] > 0]
        line["hbm_kernels_gbs"] = {k: round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) for k, v in mem}
        line["hbm_peak_gbs"] = pk["hbm"]
        if args.profile_out:
            Path(args.profile_out).write_text(json.dumps({"per_launch": prof, "summary": by}, indent=1))
        # the extras must never cost the line itself: a failure is reported inside the line
        if world == 1 and not args.no_multires and size == 1024:
            try:
                line["multires_e2e"] = multires_job(a, dev)
            except Exception as e:  # noqa: BLE001
                line["multires_e2e"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(size, args.optimizer)
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        print(json.dumps(line), flush=True)
    opt.close()
    if world > 1:
        dist.destroy_process_group()


def multires_job(a, dev):
    """BASELINE.json configs[1] as a whole job through the public driver (maua_style_b200.style.img_img_tensors): 8-bit
    RGB content + style images in pinned host memory -> preprocess -> 256 -> 512 -> 1024 with the reference's default
    iteration counts for those sizes (config.py:23-24: 500, 400, 200), L-BFGS, pastiche resident between scales ->
    deprocessed 8-bit result copied back to the host.  Wall clock around the whole call, device synchronised."""
    import copy

    import numpy as np

    from maua_style_b200 import image_ops, style

    b = copy.copy(a)
    b.image_sizes, b.num_iters, b.init, b.style_scale = [256, 512, 1024], [500, 400, 200], "content", 1.0
    rs = np.random.RandomState(0)

    def host_image(h, w, seed):  # smooth 8-bit RGB image (what a decoded photo looks like to the pipeline)
        g = torch.Generator().manual_seed(seed)
        low = torch.rand(1, 3, h // 8, w // 8, generator=g)
        img = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)[0]
        return (img * 255).clamp(0, 255).byte().permute(1, 2, 0).contiguous().pin_memory()

    content_u8, style_u8 = host_image(1024, 1024, 101), host_image(1024, 1024, 102)
    result = torch.empty(1024, 1024, 3, dtype=torch.uint8).pin_memory()

    def job():
        content = image_ops.preprocess(content_u8, dev)
        sty = image_ops.preprocess(style_u8, dev)
        outs = style.img_img_tensors(content, [sty], b)
        result.copy_(image_ops.deprocess_u8(outs[-1]), non_blocking=True)
        torch.cuda.synchronize()

    job()  # warm-up (plan core, arenas, graphs)
    t0 = time.perf_counter()
    job()
    dt = time.perf_counter() - t0
    iters = sum(b.num_iters)
    return {"workload": "multi-res 256->512->1024, L-BFGS 500/400/200 iterations, 1 style, host u8 in -> host u8 out",
            "seconds": dt, "iterations": iters, "value": iters / dt, "unit": UNIT, "images_per_min": 60.0 / dt,
            "h2d_bytes": int(content_u8.numel() + style_u8.numel()), "d2h_bytes": int(result.numel())}


def run_streams(args):
    """Experimental (batch jobs): S independent images on ONE GPU, each with its own plan, optimizer
    state, captured graph and stream, so that one image's HBM-bound phase (L-BFGS history sweeps, pools, Grams) can run
    under another image's tensor-bound convolutions.  Reports the aggregate iterations/sec next to the single-stream rate
    measured in the same process.  Single process / single GPU only."""
    from maua_style_b200 import _lib, models, optim, synthetic as O

    _lib.require_gpu()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    size, K, W, S = args.size, args.steps, args.warmup, args.streams
    tmp = tempfile.mkdtemp(prefix="maua_bench_streams_")
    ckpt = Path(tmp) / "vgg19-random.pth"
    O.save_random_checkpoint(ckpt)
    style = O.synthetic_image(size, size, seed=2).to(dev)
    jobs = []
    for k in range(S):
        a = O.reference_args(ckpt, tmp, optimizer=args.optimizer, gpu="0")
        stream = torch.cuda.Stream(dev)
        with torch.cuda.stream(stream):
            net, losses = models.load_model(a)  # a plan core is handed to one live network at a time: S cores
            optim.set_content_targets(net, O.synthetic_image(size, size, seed=1 + 10 * k, smooth=True).to(dev), a)
            optim.set_style_targets(net, [style], a)
            for m in losses:
                m.mode = "loss"
            pastiche = (O.synthetic_image(size, size, seed=4 + 10 * k) * 0.25).to(dev).contiguous()
            opt = optim.PixelOptimizer(pastiche, args.optimizer, lr=1.0, history=100)
            up = torch.zeros(net._n_slots, device=dev)
            up[net._live_slots()] = 1.0
            step = optim.GraphedIteration(net, pastiche, opt, up)
            for _ in range(args.history_prefill if args.optimizer == "lbfgs" else 3):
                step()
        stream.synchronize()
        jobs.append((stream, step, net, losses, opt, pastiche))

    def timed(active):
        for _ in range(W):
            for stream, step, *_ in active:
                with torch.cuda.stream(stream):
                    step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            for stream, step, *_ in active:  # round-robin submission: one graph launch per image and iteration
                with torch.cuda.stream(stream):
                    step()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    single = K / timed(jobs[:1])
    multi = S * K / timed(jobs)
    print(json.dumps({"metric": METRIC, "experimental": "streams", "streams": S, "value": multi, "unit": UNIT,
                      "single_stream_value": single, "speedup": multi / single, "steps": K, "warmup": W,
                      "config": workload_config(size, args.optimizer),
                      "timing": "wall clock around K round-robin graph launches per stream, device synchronised"}), flush=True)
    for _, _, _, _, opt, _ in jobs:
        opt.close()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.streams > 1:
        run_streams(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
