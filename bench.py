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
  cpu_baseline  the reference's own CPU path on the host cores, bounded sample

`--impl reference` times the UNMODIFIED reference (baseline/_ref: loss.py / models.py / optim.py as installed by
__graft_entry__.build()) driving `optim.optimize` with torch.optim.LBFGS / Adam on all host threads; only when that
directory is absent does it fall back to the oracle port (`kind: "port"`).

Extra legs of the default N = 1 line (each bounded to a few seconds): `sizes` (512^2 and 2048^2), `adam_1024`,
`covariance_2styles_1024` (BASELINE.json configs[2]), `multires_e2e` (configs[1] as a whole job), `multidevice_2048`
(configs[3], only when two GPUs are visible) and, at every N, `batch_job` (configs[4]: 8 images per GPU through
shard.stylize_images, wall clock incl. per-image set-up and D2H).
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
    ap.add_argument("--no-extras", action="store_true", help="skip the extra legs (sizes / adam / covariance / batch job / multidevice)")
    ap.add_argument("--covariance", action="store_true", help="main workload = configs[2]: covariance loss, 2 blended styles")
    ap.add_argument("--multidevice", default=None, help="run ONLY the layer-wise split leg on these devices, e.g. 0,1 (configs[3])")
    ap.add_argument("--batch-images", type=int, default=8, help="images per GPU of the batch_job leg")
    ap.add_argument("--batch-iters", type=int, default=100, help="L-BFGS iterations per image of the batch_job leg")
    ap.add_argument("--video-leg", action="store_true", help="run ONLY the vid_img frame-driver leg and print its JSON")
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


def measure_tf32_peak(dev):
    """cuBLAS TF32 matmul 8192^3 on this GPU, in this run: best single call (burst) and a 0.5 s back-to-back loop (sustained)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize(dev)
        best = 1e9
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(int(500.0 / best), 10)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record(); e1.synchronize()
        flop = 2.0 * n ** 3
        return {"burst": flop / (best * 1e-3) / 1e12, "sustained": flop * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b


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
        "weights": "He-normal random init (seed 0)", "init": "content image (--init content / the up-sampled previous scale)",
        "timing": "inputs larger than L2: 1.2 GB of activations per step",
        "sharding": "one independent image per GPU, no collective",
    }


# ---------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the UNMODIFIED reference (baseline/_ref) on the host cores
# ---------------------------------------------------------------------------------------------------------
def _reference_rate(size, optimizer, steps, warmup, budget_s):
    """it/s of the reference's own `optim.optimize` (optim.py:111-255: torch.optim.LBFGS / Adam, its own modules, CPU) on a
    bounded sample: K timed iterations after W warm-up iterations inside ONE optimize() call, timestamps taken by a
    forward pre-hook on the reference's network (one closure evaluation = one iteration, optim.py:201-238).  When a full-
    size iteration does not fit the budget the same job runs on a half / quarter-size image and the rate is scaled by the
    pixel ratio (convolution cost is linear in pixels)."""
    from baseline import ref_loader
    from maua_style_b200 import synthetic as S  # seeded inputs / checkpoint only

    if ref_loader.ref_dir() is None:
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_flush_denormal(True)
    ref = ref_loader.import_reference("stock")
    tmp = Path(tempfile.mkdtemp(prefix="maua_refarm_"))
    ckpt = tmp / "vgg19-random.pth"
    S.save_random_checkpoint(ckpt)
    rargs = ref_loader.reference_args(ref, tmp, ckpt, gpu="c", optimizer=optimizer)
    net, losses = ref.models.load_model(rargs)

    def job(sz, n_iters, stamps):
        content = S.synthetic_image(sz, sz, seed=1, smooth=True)
        style = S.synthetic_image(sz, sz, seed=2)
        init = content.clone()  # same starting point as the GPU arm (Job): the content image
        h = net.register_forward_pre_hook(lambda m, inp: stamps.append(time.perf_counter()))
        try:
            ref.optim.optimize(content, [style], init.clone(), n_iters, rargs, net, losses)
        finally:
            h.remove()

    # probe: 1 iteration at full size (2 capture forwards + 1 or 2 closure evaluations)
    stamps = []
    t0 = time.perf_counter()
    job(size, 1, stamps)
    probe = (time.perf_counter() - t0) / 3.0
    sub, scale = size, 1.0
    while probe * scale * (steps + warmup + 3) > budget_s and sub >= 256:
        sub //= 2
        scale = (sub * sub) / float(size * size)
    n_iters = steps + warmup + (1 if optimizer == "lbfgs" else 0)  # lbfgs: n evaluations, adam: n + 1 (optim.py:240)
    stamps = []
    job(sub, n_iters, stamps)
    stamps = stamps[2:]  # the two capture forwards (content, 1 style) come first
    dt = stamps[warmup + steps] - stamps[warmup]
    sample = f"{steps} iterations of the unmodified reference optim.optimize ({optimizer}, torch.optim) at {sub}x{sub} after {warmup} warm-up"
    if sub != size:
        sample += f", scaled by {scale:.4f} to {size}x{size}-equivalent (full-size probe {probe:.2f} s/iteration)"
    ref_loader.unload()
    return {"value": steps / dt * scale, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference", "sample": sample,
            "ms_per_step": dt / steps * 1e3 / scale}


def _port_rate(size, optimizer, steps, warmup, budget_s):
    """Fallback when baseline/_ref is absent: the oracle port's feval + its restated optimizer."""
    from oracle import maua_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_flush_denormal(True)
    params = O.he_init_vgg19(0)
    cfg = O.StyleConfig(content_weight=5.0, optimizer=optimizer)

    def run(sz, n):
        content = O.synthetic_image(sz, sz, seed=1, smooth=True)
        style = O.synthetic_image(sz, sz, seed=2)
        init = O.synthetic_image(sz, sz, seed=4) * 0.25
        t0 = time.perf_counter()
        O.optimize(content, [style], init, n, cfg, params)
        return time.perf_counter() - t0

    probe = run(size, 1) / 3.0
    sub, scale = size, 1.0
    while probe * scale * (steps + warmup + 3) > budget_s and sub >= 256:
        sub //= 2
        scale = (sub * sub) / float(size * size)
    t_small = run(sub, max(warmup, 1))
    t_big = run(sub, max(warmup, 1) + steps)
    dt = t_big - t_small
    sample = f"{steps} iterations of the oracle port ({optimizer}) at {sub}x{sub}"
    if sub != size:
        sample += f", scaled by {scale:.4f} to {size}x{size}-equivalent"
    return {"value": steps / dt * scale, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample,
            "ms_per_step": dt / steps * 1e3 / scale}


def cpu_rate(size, optimizer, steps, warmup, budget_s):
    import contextlib

    r = None
    try:
        # the reference's progress bar writes to sys.stdout (optim.py:19); stdout carries the JSON line only
        with contextlib.redirect_stdout(sys.stderr):
            r = _reference_rate(size, optimizer, steps, warmup, budget_s)
    except Exception as e:  # noqa: BLE001  (a broken reference install must not cost the line)
        sys.stderr.write(f"reference arm: unmodified reference failed ({type(e).__name__}: {e}); using the oracle port\n")
    return r if r is not None else _port_rate(size, optimizer, steps, warmup, budget_s)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_rate(args.size, args.optimizer, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.size, args.optimizer),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(size, optimizer):
    """Bounded sample (about 10-30 s of CPU work) of the same step on the host cores."""
    r = cpu_rate(size, optimizer, steps=3, warmup=1, budget_s=25.0)
    return {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}


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


LBFGS_LAUNCHES = 4  # multi-dot pass, partial reduce, scalar recurrences, direction + update (maua_style_b200/csrc/lbfgs.cu)


class Job:
    """One image being optimised exactly as optim.optimize drives it (2 eager rounds, then CUDA-graph replays)."""

    def __init__(self, size, optimizer, dev, local_rank, rank=0, covariance=False, history_prefill=100, multidevice=None,
                 arch="vgg19"):
        from maua_style_b200 import models, optim, synthetic as O

        self.tmp = tempfile.mkdtemp(prefix=f"maua_bench_{rank}_")
        over = dict(optimizer=optimizer, gpu=str(local_rank))
        if arch == "prune":  # models.py:136, :249-258: the channel-pruned VGG-16 (stock schedule above 2656 px)
            ckpt = Path(self.tmp) / "vgg16-prune-random.pth"
            O.save_random_checkpoint(ckpt, channels=models.channel_list["VGG-16p"])
        elif arch == "nin":  # models.py:74-113; layer lists of config/scaling-img.json:32-47 (stock schedule above 4096 px)
            ckpt = Path(self.tmp) / "nin-random.pth"
            O.save_random_checkpoint(ckpt, channels=models.NIN_LAYERS)
            over.update(style_layers="relu1,relu3,relu5,relu7,relu9,relu11", content_layers="relu8")
        else:
            ckpt = Path(self.tmp) / "vgg19-random.pth"
            O.save_random_checkpoint(ckpt)
        if covariance:  # BASELINE.json configs[2]: covariance loss, 2 blended styles of different aspect (area-matched)
            over.update(use_covariance=True, style_blend_weights="3,1")
        if multidevice:
            over.update(gpu=multidevice, multidevice=True, multidevice_strategy="5")
        self.args = a = O.reference_args(ckpt, self.tmp, **over)
        self.net, self.losses = models.load_model(a)
        content = O.synthetic_image(size, size, seed=1 + 10 * rank, smooth=True)
        if covariance:
            h2 = int(size * 0.875) // 2 * 2
            w2 = (size * size // h2) // 2 * 2
            styles = [O.synthetic_image(h2, w2, seed=2), O.synthetic_image(w2, h2, seed=3, smooth=True)]
        else:
            styles = [O.synthetic_image(size, size, seed=2)]
        # the pastiche starts from the content image (style.py:56-57 `--init content`; at the later scales of a multi-resolution job
        # it is the up-sampled previous result, style.py:64-66): an image-like starting point.  A random starting point makes the
        # line-search-free L-BFGS of the reference reject every curvature pair at <= 512^2 with random-init weights (y.s <= 1e-10,
        # torch's gate), which would time the loop at history 1 instead of 100 -- the history length reached is reported
        init = content.clone()
        optim.set_content_targets(self.net, content.to(dev), a)
        optim.set_style_targets(self.net, [s.to(dev) for s in styles], a)
        for m in self.losses:
            m.mode = "loss"
        self.pastiche = init.to(dev).contiguous()
        self.opt = optim.PixelOptimizer(self.pastiche, optimizer, lr=1.0, history=100)
        self.up = torch.zeros(self.net._n_slots, device=dev)
        self.up[self.net._live_slots()] = 1.0
        self.step = optim.GraphedIteration(self.net, self.pastiche, self.opt, self.up)
        self.optimizer = optimizer
        # set-up: fill the L-BFGS history so the timed steps run at full history (not part of warm-up or timing)
        for _ in range(history_prefill if optimizer == "lbfgs" else 3):
            self.step()
        torch.cuda.synchronize(dev)

    def lbfgs_state(self):
        """(iterations, history length, halted flag) of the device L-BFGS, or None for Adam."""
        import ctypes as C

        from maua_style_b200 import _lib

        if self.optimizer != "lbfgs":
            return None
        n, h, halted = C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.load().maua_lbfgs_query(self.opt._state, C.byref(n), C.byref(h), C.byref(halted), _lib.stream_ptr()))
        return {"iterations": n.value, "history_len": h.value, "halted": bool(halted.value),
                "pastiche_finite": bool(torch.isfinite(self.pastiche).all())}

    def launches_per_step(self):
        f, b = self.net.last_launches()
        return f + b + (LBFGS_LAUNCHES if self.optimizer == "lbfgs" else 2)  # adam: step counter + update

    def timed(self, K, W, barrier=None):
        """ms for K steps, CUDA events on the launching stream, synchronised on both sides."""
        for _ in range(W):
            self.step()
        (barrier or torch.cuda.synchronize)()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            self.step()
        e1.record()
        (barrier or torch.cuda.synchronize)()
        return e0.elapsed_time(e1)

    def conv_profile(self, reps=3):
        """Per-launch CUDA-event records of one feval (plan profile mode), last of `reps` evaluations."""
        self.net.set_profile(True)
        prof = None
        for _ in range(reps):
            self.net._forward_plan(self.pastiche, keep=True)
            self.net._backward_plan(self.up)
            prof = self.net.profile()
        self.net.set_profile(False)
        return prof

    def close(self):
        self.opt.close()


def conv_summary(prof):
    conv = [r for r in prof if r["name"] in ("conv_fwd", "conv_dgrad")]
    ms = sum(r["ms"] for r in conv)
    flops = sum(r["flops"] for r in conv)
    return conv, ms, flops, sum(r["ms"] for r in prof)


def side_leg(size, optimizer, dev, K, W, pk, covariance=False, prefill=100, arch="vgg19"):
    """A bounded extra measurement of another configuration: it/s (device-resident, CUDA events) + its conv roofline."""
    job = Job(size, optimizer, dev, dev.index, covariance=covariance, history_prefill=prefill, arch=arch)
    ms = job.timed(K, W)
    conv, conv_ms, conv_flops, feval_ms = conv_summary(job.conv_profile())
    tf = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    out = {"size": size, "optimizer": optimizer, "value": K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "steps": K,
           "warmup": W, "conv_tflops": tf, "conv_frac_of_tf32_burst": tf / (pk["bf16_burst"] / 2.0),
           "feval_ms": feval_ms, "gpu_launches_per_step": job.launches_per_step(), "cuda_graph": job.step.graph is not None}
    if optimizer == "lbfgs":
        out["lbfgs_state"] = job.lbfgs_state()
    if covariance:
        out["styles"] = 2
        out["loss"] = "covariance (--use_covariance), blend 3:1"
    if arch != "vgg19":
        out["model"] = {"prune": "channel-pruned VGG-16 (zero-padded to tileable channel counts)",
                        "nin": "NIN (1x1 / 3x3 / 5x5 layers: conv_tc_kernel; 11x11/4 image layer: im2col + pointwise tcgen05 GEMM; 3x3/2 ceil pools)"}[arch]
        out["kernel_breakdown_ms"] = {}
        for r in job.conv_profile():
            out["kernel_breakdown_ms"][r["name"]] = round(out["kernel_breakdown_ms"].get(r["name"], 0.0) + r["ms"], 4)
    job.close()
    del job
    torch.cuda.empty_cache()
    return out


def video_window_leg(dev, size, frames, K, W):
    """One img_vid window (optim.py:113-170): `frames` pastiche frames as ONE batch through the network -- per-frame static
    Grams, the [frames*C, frames*C] dynamic Gram (one SYRK over the side-by-side features), Adam.  it/s counts window
    iterations (each is `frames` forward + backward passes)."""
    from maua_style_b200 import models, optim, synthetic as O

    tmp = tempfile.mkdtemp(prefix="maua_bench_vid_")
    ckpt = Path(tmp) / "vgg19-random.pth"
    O.save_random_checkpoint(ckpt)
    a = O.reference_args(ckpt, tmp, optimizer="adam", gpu=str(dev.index), transfer_type="img_vid", gram_frame_window=frames,
                         avg_frame_window=-1)
    net, losses = models.load_model(a)
    content = O.synthetic_image(size, size, seed=1, smooth=True).to(dev)
    video = torch.cat([O.synthetic_image(size, size, seed=20 + f, smooth=(f % 2 == 0)) for f in range(frames + 1)]).to(dev)
    optim.set_content_targets(net, content, a)
    optim.set_style_video_targets(net, [video], a)
    for m in losses:
        m.mode = "loss"
    pastiche = torch.cat([O.synthetic_image(size, size, seed=40 + f) * 0.25 for f in range(frames)]).to(dev).contiguous()
    opt = optim.PixelOptimizer(pastiche, "adam", lr=1.0)
    up = torch.zeros(net._n_slots, device=dev)
    up[net._live_slots()] = 1.0
    step = optim.GraphedIteration(net, pastiche, opt, up)
    for _ in range(W + 3):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    opt.close()
    return {"workload": f"one img_vid window: {frames} frames of {size}x{size}, style video of {frames + 1} frames, static + dynamic "
                        f"([{frames}*C]^2) Gram losses, Adam", "value": K / (ms * 1e-3), "unit": "window it/s",
            "frame_evals_per_sec": frames * K / (ms * 1e-3), "ms_per_step": ms / K, "steps": K, "warmup": W,
            "cuda_graph": step.graph is not None}


def video_frames_leg(dev, size=1024, frames=8, iters=100, passes=2):
    """BASELINE.json configs[4] in its frame-chunk form, per GPU: `frames` consecutive frames through the public video driver
    (maua_style_b200.style.vid_img_tensors; style.py:145-300): one scale, `passes` passes of iters // passes L-BFGS iterations per
    frame, the previous result warped along the flow as the temporal target from the second pass on, results handed on as
    8-bit frames.  Flow fields / reliability maps are synthetic inputs (flow estimation is not part of the path).  Wall clock
    around the whole call, device synchronised on both sides; second (warm) run reported."""
    from maua_style_b200 import style, synthetic as O

    tmp = tempfile.mkdtemp(prefix="maua_bench_frames_")
    ckpt = Path(tmp) / "vgg19-random.pth"
    O.save_random_checkpoint(ckpt)
    clip = [O.synthetic_image(size, size, seed=100 + i, smooth=True).to(dev) for i in range(frames)]
    styles = [O.synthetic_image(size, size, seed=2).to(dev)]
    g = torch.Generator().manual_seed(5)
    flow = torch.randn(size // 4, size // 4, 2, generator=g) * 0.002  # normalised + blurred field (style.read_flo's output)
    rel = (torch.rand(1, 1, size // 4, size // 4, generator=g) > 0.1).float().to(dev)
    flows = lambda d, i, j: (flow if d == "forward" else -flow, rel)

    def job():
        a = O.reference_args(ckpt, tmp, transfer_type="vid_img", optimizer="lbfgs", gpu=str(dev.index), image_sizes=[size],
                             num_iters=[iters], passes_per_scale=passes, init="content", temporal_blend=0.5, loop=False,
                             style_scale=1.0, match_histograms=False)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        store = style.vid_img_tensors(clip, styles, a, flows)
        torch.cuda.synchronize(dev)
        return time.perf_counter() - t0, len(store)

    job()
    dt, n = job()
    evals = frames * passes * (iters // passes)
    return {"workload": f"vid_img driver: {frames} consecutive {size}x{size} frames, 1 scale, {passes} passes x {iters // passes} L-BFGS "
                        "iterations per frame, warp + flow-weighted temporal loss from pass 2, synthetic frames / flows",
            "frames": frames, "seconds": dt, "frames_per_min": 60.0 * frames / dt, "iterations": evals, "value": evals / dt,
            "unit": UNIT, "results": n, "timing": "wall clock around style.vid_img_tensors, device synchronised on both sides"}


def batch_job(dev, info, images_per_gpu, iters, size=1024):
    """BASELINE.json configs[4] per GPU: `images_per_gpu` independent content images at 1024^2 through the public sharded
    runner (shard.stylize_images: one network per rank, style targets captured once, per-image content capture +
    optimisation + device->host copy of the result).  Wall clock around the whole call; with N ranks the job is
    N * images_per_gpu images (8 x 8 = the 64-image job at N = 8)."""
    from maua_style_b200 import shard, synthetic as O

    tmp = tempfile.mkdtemp(prefix=f"maua_batch_{info.rank}_")
    ckpt = Path(tmp) / "vgg19-random.pth"
    O.save_random_checkpoint(ckpt)
    a = O.reference_args(ckpt, tmp, optimizer="lbfgs", gpu=str(info.local_rank))
    n = images_per_gpu * info.world
    style = O.synthetic_image(size, size, seed=2)
    mine = set(shard.partition_round_robin(n, info.world, info.rank))
    # every rank only materialises its own images (pinned host memory: what a decoded batch looks like)
    contents = [O.synthetic_image(size, size, seed=100 + i, smooth=True).pin_memory() if i in mine else None for i in range(n)]
    inits = contents
    first = contents[min(mine)]
    shard.stylize_images([first], [style], [first], 2, a, info=shard.RankInfo(0, 1, info.local_rank))  # warm-up: plan core, arena
    shard.barrier()
    t0 = time.perf_counter()
    out = shard.stylize_images(contents, [style], inits, iters, a, info=info)
    dt = shard.max_over_ranks(time.perf_counter() - t0)
    assert len(out) == len(mine) and all(v.device.type == "cpu" for v in out.values())
    return {"workload": f"{n} independent {size}x{size} content images ({images_per_gpu} per GPU), {iters} L-BFGS iterations each, "
                        "shard.stylize_images (per-image content capture, optimisation, D2H of the result)",
            "images": n, "iterations_per_image": iters, "seconds": dt, "images_per_min": 60.0 * n / dt,
            "value": n * iters / dt, "unit": UNIT, "timing": "wall clock, max over ranks, barrier on both sides"}


def multidevice_leg(devices, size, K, W):
    """BASELINE.json configs[3]: one image, the reference's layer-wise split (models.py:503-566) over `devices`, Adam (the
    reference switches to Adam above 1456 px, config/scaling-img.json).  The hand-over tensors are stored into the
    consuming GPU's memory by the producing kernels over NVLink.  Reported next to the same image on one GPU."""
    dev0 = torch.device("cuda", int(devices.split(",")[0]))
    out = {"devices": devices, "size": size, "optimizer": "adam"}
    for tag, md in (("single_gpu", None), ("split", devices)):
        job = Job(size, "adam", dev0, dev0.index, multidevice=md)
        for _ in range(W):
            job.step()
        for d in range(torch.cuda.device_count()):
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for _ in range(K):
            job.step()
        for d in range(torch.cuda.device_count()):
            torch.cuda.synchronize(d)
        dt = time.perf_counter() - t0
        rec = {"value": K / dt, "unit": UNIT, "ms_per_step": dt / K * 1e3, "steps": K}
        if md:
            stages = job.net._stages
            rec["stages"] = [{"device": str(st["device"]), "entries": [st["begin"], st["end"]]} for st in stages]
            rec["handoff_bytes_per_step"] = int(sum(2 * st["x_in"].numel() * 4 for st in stages if st["x_in"] is not None))
            rec["handoff"] = "forward activation + backward gradient of the stage boundary, peer stores over NVLink"
        out[tag] = rec
        job.close()
        del job
        torch.cuda.empty_cache()
    out["timing"] = "wall clock around K eager iterations (several devices: no single-stream events), all devices synchronised"
    return out


def run_ours(args):
    import torch.distributed as dist

    from maua_style_b200 import _lib, optim, shard  # the product path never touches oracle/

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.require_gpu()
    info = shard.RankInfo(rank, world, local_rank)

    if args.multidevice:
        print(json.dumps({"metric": METRIC, "leg": "multidevice", **multidevice_leg(args.multidevice, args.size, args.steps, args.warmup)}),
              flush=True)
        return

    if args.video_leg:
        print(json.dumps({"metric": METRIC, "leg": "vid_img_frames", **video_frames_leg(dev, args.size)}), flush=True)
        return

    size, K, W = args.size, args.steps, args.warmup
    sampler = ClockSampler(local_rank)
    sampler.start()
    job = Job(size, args.optimizer, dev, local_rank, rank=rank, covariance=args.covariance, history_prefill=args.history_prefill)
    net, pastiche, step = job.net, job.pastiche, job.step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    t_begin = time.time()
    ms_local = job.timed(K, W, barrier)
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end)
    ms = torch.tensor([ms_local], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * K / (ms_total / 1e3)
    gpu_launches = K * job.launches_per_step()  # kernels executed (as nodes of the replayed graph when graphs are on)

    # ---- end to end with host buffers: every step's input image comes from pinned host memory (H2D) and its result
    # (updated image + total loss) goes back to pinned host memory (D2H); copies are double-buffered on their own streams
    # (optim.HostPipelinedIteration) so they overlap the neighbouring steps' kernels ----
    pipe = optim.HostPipelinedIteration(step)
    host_in = [pastiche.detach().cpu().pin_memory(), (pastiche.detach().cpu() * 0.5).pin_memory()]
    for i in range(min(W, 3)):
        pipe.submit(host_in[i % 2])
    pipe.drain()
    barrier()
    t0 = time.perf_counter()
    got = 0
    for i in range(K):
        r = pipe.submit(host_in[i % 2])
        got += r is not None
    r = pipe.drain()
    got += r is not None
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    assert got == K and torch.isfinite(r[1]).all()
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * K / float(dt.item())

    cfg = workload_config(size, args.optimizer)
    if args.covariance:
        cfg["workload"] = (f"VGG-19 covariance style transfer {size}x{size} (--use_covariance), 2 blended styles 3:1, content relu4_2 + "
                           f"style relu1_1..relu5_1 + TV, {args.optimizer} (BASELINE.json configs[2])")
        cfg["styles"] = 2
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32", "dtype_detail": "TF32 tensor-core operands, fp32 accumulate, fp32 storage", "data": "synthetic",
        "config": cfg, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.bytes_in, "d2h_bytes_per_step": pipe.bytes_out,
                "how": "optim.HostPipelinedIteration: per step H2D of the input image from pinned memory, one iteration, D2H of the "
                       "updated image + loss; copies double-buffered on separate streams; wall clock, max over ranks"},
        "gpu_launches": gpu_launches,
        "cuda_graph": bool(step.graph is not None),
        "lbfgs_state": job.lbfgs_state(),
    }

    # ---- configs[4] at this N: images_per_gpu x N images through the sharded runner (all ranks take part) ----
    if not args.no_extras and size == 1024 and not args.covariance:
        try:
            bj = batch_job(dev, info, args.batch_images, args.batch_iters)
        except Exception as e:  # noqa: BLE001
            bj = {"error": f"{type(e).__name__}: {e}"[:300]}
        line["batch_job"] = bj

    if rank == 0:
        # ---- roofline of the dominant kernel family, measured live with per-launch CUDA events ----
        pk = peaks()
        prof = job.conv_profile()
        conv, conv_ms, conv_flops, iter_ms = conv_summary(prof)
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        tf32_burst, tf32_sustained = pk["bf16_burst"] / 2.0, pk["bf16"] / 2.0
        try:
            cublas = measure_tf32_peak(dev)
        except Exception as e:  # noqa: BLE001
            cublas = {"error": f"{type(e).__name__}: {e}"[:200]}
        line["roofline"] = {
            "kernel": f"conv_tc_kernel (tcgen05 implicit-GEMM conv3x3, forward + dgrad, {len(conv)} launches/iteration)",
            "bound": "tensor", "achieved": achieved, "peak": tf32_burst, "unit": "TFLOP/s", "frac": achieved / tf32_burst,
            "peak_source": f"{pk['source']}: bf16_tflops (burst) {pk['bf16_burst']:.1f} / 2 -- TF32 MMAs issue at half the bf16 rate; "
                           "burst because the launches are timed one by one with CUDA events",
            "frac_vs_sustained_peak": achieved / tf32_sustained, "sustained_peak": tf32_sustained,
            "cublas_tf32_8192_measured_in_this_run": cublas,
            "frac_of_bf16_peak": achieved / pk["bf16_burst"],
            "traffic": conv_traffic_per_launch(size),
            "traffic_source": "profiles/conv_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per conv_tc launch, ncu --set full "
                              "capture of this command, committed; not measurable inside an un-profiled run)",
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
        a = job.args
        job.close()
        del job, pipe, step, net
        torch.cuda.empty_cache()
        # the extras must never cost the line itself: a failure is reported inside the line
        if world == 1 and size == 1024 and not args.covariance:
            def guarded(fn, *fa, **kw):
                try:
                    return fn(*fa, **kw)
                except Exception as e:  # noqa: BLE001
                    return {"error": f"{type(e).__name__}: {e}"[:300]}

            if not args.no_multires:
                line["multires_e2e"] = guarded(multires_job, a, dev)
            if not args.no_extras:
                line["sizes"] = {"512": guarded(side_leg, 512, args.optimizer, dev, 40, 5, pk),
                                 "2048": guarded(side_leg, 2048, args.optimizer, dev, 10, 3, pk)}
                line["adam_1024"] = guarded(side_leg, 1024, "adam", dev, 30, 5, pk)
                line["covariance_2styles_1024"] = guarded(side_leg, 1024, args.optimizer, dev, 30, 5, pk, covariance=True)
                # the other backbones of the stock schedule (config/scaling-img.json: "prune" above 2656 px, "nin" above 4096 px)
                line["pruned_vgg16_2048_adam"] = guarded(side_leg, 2048, "adam", dev, 10, 3, pk, arch="prune")
                line["nin_4096_adam"] = guarded(side_leg, 4096, "adam", dev, 5, 2, pk, arch="nin")
                line["img_vid_window"] = guarded(video_window_leg, dev, 512, 4, 10, 3)
                line["vid_img_frames"] = guarded(video_frames_leg, dev)
                if torch.cuda.device_count() >= 2:
                    line["multidevice_2048"] = guarded(multidevice_leg, "0,1", 2048, 10, 3)
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(size, args.optimizer)
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        print(json.dumps(line), flush=True)
    else:
        job.close()
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
