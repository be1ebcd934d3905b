"""Sharding independent style-transfer jobs over the GPUs of one box -- one process per GPU, NO data-path collective.

The reference has no distributed backend at all: `style.py` optimises one content image
(img_img, style.py:22-73) or walks the frames of a video one by one (vid_img, style.py:176-290).  Every content image
is an independent optimisation, and within a video pass frame n only needs frame n-1's output (style.py:273-286), so
the natural partition over an 8 x B200 box is

  * images  : round-robin  (job i -> rank i % world), or
  * frames  : contiguous chunks (rank r owns frames [r*ceil(n/world), ...)); chunk heads take their init image from
              the previous pass / scale exactly like the reference's resume path does (style.py:232-271).

Weights (52 MB) are replicated and the style targets (<= 2.4 MB) are recomputed per rank (1 forward), so independent images
exchange nothing, and neither do video-frame chunks styled as clips of their own (`stylize_video_chunks`).  A video whose chunk
seams stay temporally coupled (`stylize_video`) has one exchange step per pass: the 8-bit frames every rank
produced are all-gathered (`exchange_frames`), because chunk heads start from their predecessor's previous-pass result.
Otherwise torch.distributed is only used for the control plane: the rendezvous, a barrier around timed regions
and the max-over-ranks of the elapsed time (bench.py).  Works with the `gloo` backend on CPU (tests) and `nccl` on GPUs.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Iterable, List, Optional, Sequence, Tuple


@dataclass(frozen=True)
class RankInfo:
    rank: int
    world: int
    local_rank: int

    @staticmethod
    def from_env() -> "RankInfo":
        """RANK / WORLD_SIZE / LOCAL_RANK as set by torchrun (single process: 0 / 1 / 0)."""
        return RankInfo(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                        int(os.environ.get("LOCAL_RANK", "0")))


def partition_round_robin(n_items: int, world: int, rank: int) -> List[int]:
    """Independent images: job i runs on rank i % world."""
    _check(n_items, world, rank)
    return list(range(rank, n_items, world))


def partition_contiguous(n_items: int, world: int, rank: int) -> List[int]:
    """Video frames: rank r owns one contiguous chunk; chunk sizes differ by at most one frame and earlier ranks
    get the longer chunks (so frame order == rank order)."""
    _check(n_items, world, rank)
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def chunk_bounds(n_items: int, world: int) -> List[Tuple[int, int]]:
    """[start, end) of every rank's contiguous chunk."""
    out = []
    for r in range(world):
        idx = partition_contiguous(n_items, world, r)
        out.append((idx[0], idx[-1] + 1) if idx else (0, 0))
    return out


def _check(n_items: int, world: int, rank: int) -> None:
    if n_items < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad partition request: n_items={n_items} world={world} rank={rank}")


def init_process_group(backend: Optional[str] = None, info: Optional[RankInfo] = None):
    """Rendezvous for the control plane.  Returns the RankInfo; a no-op for a single process."""
    import torch
    import torch.distributed as dist

    info = info or RankInfo.from_env()
    if info.world == 1 or dist.is_initialized():
        return info
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend == "nccl":
        torch.cuda.set_device(info.local_rank)
        dist.init_process_group(backend, rank=info.rank, world_size=info.world,
                                device_id=torch.device("cuda", info.local_rank))
    else:
        dist.init_process_group(backend, rank=info.rank, world_size=info.world)
    return info


def barrier() -> None:
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value: float) -> float:
    """Elapsed time of a sharded job = the slowest rank's time (control plane only, one scalar)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def run_sharded(jobs: Sequence, fn: Callable, info: Optional[RankInfo] = None, contiguous: bool = False,
                gather: bool = False):
    """Run fn(job_index, job) for this rank's share of `jobs`.

    Returns {job_index: result} for the local jobs, or -- with gather=True -- the merged dict of every rank's results
    on every rank (all_gather_object: control plane; use it for small metadata such as output paths or loss values,
    never for images).
    """
    import torch.distributed as dist

    info = info or RankInfo.from_env()
    part = partition_contiguous if contiguous else partition_round_robin
    mine = part(len(jobs), info.world, info.rank)
    local = {}
    for i in mine:
        local[i] = fn(i, jobs[i])
    if not gather or info.world == 1 or not (dist.is_available() and dist.is_initialized()):
        return local
    parts = [None] * info.world
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return merged


def stylize_images(contents: Sequence, styles: Sequence, inits: Sequence, num_iters: int, args,
                   info: Optional[RankInfo] = None, net=None, losses=None):
    """BASELINE.json config 5: a batch of independent content images, sharded one stream per GPU.

    Every rank builds its own plan on cuda:<local_rank> (weights replicated), re-uses it -- and the captured style
    targets -- for all of its images, and returns {image_index: stylised CPU tensor} for the images it owns.
    """
    import copy

    import torch

    from . import models, optim

    info = info or RankInfo.from_env()
    a = copy.copy(args)
    a.gpu = str(info.local_rank)
    if net is None or losses is None:
        net, losses = models.load_model(a)
    mine = partition_round_robin(len(contents), info.world, info.rank)
    out = {}
    # One network serves all of this rank's images, but every image is an independent img_img job: the reference builds a
    # fresh network per optimize() call (optim.py:127-129), so `--normalize_weights` (optim.py:176-178, which divides the
    # module strengths in place) applies exactly once per image.  Restore the strengths between images.
    mods = net.content_losses + net.style_losses + net.temporal_losses
    base = [m.strength for m in mods]
    for i in mine:
        out[i] = optim.optimize(contents[i], styles, inits[i], num_iters, a, net, losses)
        for m, s0 in zip(mods, base):
            m.strength = s0
    torch.cuda.synchronize()
    return out


def exchange_frames(fresh: dict, info: Optional[RankInfo] = None) -> dict:
    """The one exchange step of a sharded video job: after a pass every rank holds the 8-bit frames it styled,
    {(size, pass, frame index): uint8 [H,W,3]}, and the next pass needs its neighbours' (the frame in front of a chunk head, and
    for the file-level driver the whole pass).  One all_gather of the padded per-rank stacks (3 bytes per pixel and frame: 151 MB
    for 48 frames of 1024^2 -- against minutes of optimisation per pass) + the keys as objects.  Returns all ranks' frames."""
    import torch
    import torch.distributed as dist

    info = info or RankInfo.from_env()
    if info.world == 1 or not (dist.is_available() and dist.is_initialized()):
        return dict(fresh)
    keys = sorted(fresh)
    shape = tuple(fresh[keys[0]].shape) if keys else None
    metas = [None] * info.world
    dist.all_gather_object(metas, (keys, shape))
    shapes = {m[1] for m in metas if m[1] is not None}
    if not shapes:
        return {}
    if len(shapes) != 1:
        raise RuntimeError(f"ranks disagree on the frame size of this pass: {sorted(shapes)}")
    shape = shapes.pop()
    most = max(len(m[0]) for m in metas)
    dev = fresh[keys[0]].device if keys else (torch.device("cuda", torch.cuda.current_device())
                                              if dist.get_backend() == "nccl" else torch.device("cpu"))
    buf = torch.zeros((most,) + shape, dtype=torch.uint8, device=dev)
    for i, k in enumerate(keys):
        buf[i] = fresh[k]
    parts = [torch.empty_like(buf) for _ in range(info.world)]
    dist.all_gather(parts, buf)
    merged = {}
    for (ks, _), part in zip(metas, parts):
        for i, k in enumerate(ks):
            merged[tuple(k)] = part[i].clone()
    return merged


def stylize_video(frames: Sequence, styles: Sequence, args, flows: Callable, info: Optional[RankInfo] = None,
                  on_frame: Optional[Callable] = None) -> dict:
    """BASELINE.json config 5, video form: the frames of one clip in contiguous chunks, one chunk per GPU.
    Every rank walks the same scales x passes schedule (style.vid_img_tensors) on its own chunk; after each pass the ranks
    exchange the 8-bit frames they produced (`exchange_frames`), so chunk heads start from their predecessor's result of the
    previous pass / scale -- what the reference's resume path does with the PNGs of an interrupted run (style.py:186-188,
    :229-271).  Within a chunk results are carried from frame to frame as in the unsharded driver; only the first frame of a
    chunk differs from a single-GPU run (it sees its predecessor one pass late).  Returns every frame of every pass on every
    rank: {(size, pass, frame index): uint8 [H,W,3]}."""
    import copy

    from . import style

    info = info or RankInfo.from_env()
    a = copy.copy(args)
    a.gpu = str(info.local_rank)
    owned = partition_contiguous(len(frames), info.world, info.rank)
    return style.vid_img_tensors(frames, styles, a, flows, on_frame=on_frame, owned=owned,
                                 exchange=lambda fresh: exchange_frames(fresh, info))


def stylize_video_chunks(frames: Sequence, styles: Sequence, args, flows: Callable, info: Optional[RankInfo] = None,
                         on_frame: Optional[Callable] = None) -> dict:
    """BASELINE.json config 5 as it is worded -- "video-frame chunks ... one per GPU, with no collective": the clip is cut into
    contiguous chunks and every rank styles ITS chunk as a clip of its own (style.vid_img_tensors on the sub-list: the pair that
    closes a pass goes from the chunk's last frame back to its first, style.py:195-197), so no frame ever crosses a GPU boundary
    and nothing is exchanged.  `flows(direction, i, j)` is called with indices into `frames` and must also know the closing pair
    of every chunk.  Returns this rank's frames only: {(size, pass, frame index in `frames`): uint8 [H,W,3]}.  Chunk seams are
    not temporally coupled; `stylize_video` is the coupled variant (one all-gather of 8-bit frames per pass)."""
    import copy

    from . import style

    info = info or RankInfo.from_env()
    a = copy.copy(args)
    a.gpu = str(info.local_rank)
    idx = partition_contiguous(len(frames), info.world, info.rank)
    if len(idx) < 2:
        raise ValueError(f"rank {info.rank} would get {len(idx)} frame(s): a chunk needs at least two")
    sub = [frames[i] for i in idx]
    cb = None if on_frame is None else (lambda size, p, f, u8: on_frame(size, p, idx[f], u8))
    store = style.vid_img_tensors(sub, styles, a, lambda d, i, j: flows(d, idx[i], idx[j]), on_frame=cb)
    return {(size, p, idx[f]): v for (size, p, f), v in store.items()}
