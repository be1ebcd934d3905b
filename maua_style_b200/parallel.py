"""Layer-wise multidevice split with the reference's interface (reference models.py:503-566: `ModelParallel`,
`setup_multi_device`), for single images that are too large for one device (BASELINE.json config 4).

The reference chunks its `nn.Sequential` after the module indices listed in `--multidevice_strategy` (default "5"), puts
chunk i on device i of `--gpu` and hops activations with `.to(device)` (a stream-ordered `cudaMemcpyPeerAsync`; autograd
mirrors it in the backward pass).  Here every chunk is one stage plan (`maua_plan_create_stage`) and there is NO copy:
the hand-over tensor is allocated on the CONSUMING device and the producing stage's last kernel (conv epilogue or pool)
stores into it directly, tile by tile, over NVLink peer access while the rest of the grid is still computing; the
gradient flows back the same way (the next stage's first dgrad stores into the previous device's memory).  CUDA events
order the per-device streams.

Split points.  Our conv + bias + ReLU (+ the loss modules tapped there) is one fused entry, and a stage must begin
with a conv, so a requested module index is moved forward to the end of the entry that contains it and past a
directly following pool -- which is also where the hand-over is cheapest (the tensor after a pool
is 4x smaller).  E.g. the reference default "5" (after conv1_2, before its ReLU: a 1.07 GB tensor at 2048^2) becomes
"after pool1" (268 MB).
"""
from __future__ import annotations

from typing import List

import torch


def module_index_map(entries: List[int], taps, has_tv: bool, has_temporal: bool) -> List[int]:
    """Index of the LAST reference module (in the nn.Sequential `load_model` would build, models.py:369-436) that
    belongs to each of our fused entries."""
    idx = int(has_tv) + int(has_temporal) - 1  # modules before conv1_1
    per_relu = {}
    for ridx, _ in taps:
        per_relu[ridx] = per_relu.get(ridx, 0) + 1
    last, conv = [], -1
    for c in entries:
        if c == 0:
            idx += 1                # pool
        else:
            conv += 1
            idx += 2                # conv, relu
            idx += per_relu.get(conv, 0)  # loss modules spliced after this relu
        last.append(idx)
    return last


def stage_bounds(entries: List[int], strategy: str, n_devices: int, taps, has_tv: bool, has_temporal: bool) -> List[int]:
    """Entry ranges [b0=0, b1, ..., len(entries)] for `--multidevice_strategy` (comma separated module indices)."""
    splits = [int(x) for x in str(strategy).split(",") if str(x).strip() != ""]
    # models.py:541-543
    assert n_devices - 1 == len(splits), \
        "The number of -multidevice_strategy layer indices must be equal to the number of -gpu devices minus 1."
    last = module_index_map(entries, taps, has_tv, has_temporal)
    bounds = [0]
    for sp in splits:
        e = next((i for i, l in enumerate(last) if l >= sp), len(entries) - 1)  # entry containing module `sp`
        b = e + 1
        if b < len(entries) and entries[b] == 0:  # a stage must begin with a conv: keep the pool with its producer
            b += 1
        b = max(b, bounds[-1] + 1)
        if b >= len(entries):
            raise ValueError(f"multidevice_strategy {strategy!r}: split index {sp} leaves no layers for the next device "
                             f"(the truncated network has {last[-1] + 1} modules)")
        bounds.append(b)
    bounds.append(len(entries))
    return bounds


def setup_multi_device(entries, params, args, taps, tv_mod, temporal_mod, content_losses, style_losses, tv_losses,
                       temporal_losses, norm_channels=None, conv_kinds=None, pool_kind=0):
    """models.py:537-566 + :440-441: returns (net, losses) with `net` spanning the devices of `--gpu`."""
    from .models import build_net

    gpus = [g for g in str(args.gpu).split(",") if g != ""]
    if any(g.lower() == "c" for g in gpus):
        raise RuntimeError("maua_style_b200 has no CPU path: --gpu entries must be CUDA device indices")
    devices = [torch.device("cuda", int(g)) for g in gpus]
    if len(devices) < 2:
        raise ValueError("--multidevice needs at least two devices in --gpu")
    if torch.cuda.device_count() <= max(d.index for d in devices):
        raise RuntimeError(f"--gpu {args.gpu}: only {torch.cuda.device_count()} CUDA device(s) are visible")
    bounds = stage_bounds(entries, getattr(args, "multidevice_strategy", "5"), len(devices), taps, tv_mod is not None,
                          temporal_mod is not None)
    if getattr(args, "verbose", False):
        for k, d in enumerate(devices):
            print(f"device {d}: entries [{bounds[k]}, {bounds[k + 1]})")
    net = build_net(args, entries, params, taps, tv_mod, temporal_mod, devices[0], stage_bounds=bounds, devices=devices,
                    norm_channels=norm_channels, conv_kinds=conv_kinds, pool_kind=pool_kind)
    net.content_losses = content_losses
    net.style_losses = style_losses
    net.tv_losses = tv_losses
    net.temporal_losses = temporal_losses
    return net, content_losses + style_losses + tv_losses + temporal_losses
