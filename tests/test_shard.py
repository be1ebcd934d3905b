"""Host-side sharding logic (maua_style_b200/shard.py): partitions, and the world_size-2 control plane over gloo.

No GPU and no compute calls here: the jobs are stand-ins; the data path of the real thing has no collective at all.
"""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maua_style_b200 import shard

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("n,world", [(0, 1), (1, 1), (7, 2), (8, 8), (64, 8), (5, 8), (13, 4)])
def test_partitions_cover_every_item_exactly_once(n, world):
    for part in (shard.partition_round_robin, shard.partition_contiguous):
        seen = []
        sizes = []
        for r in range(world):
            idx = part(n, world, r)
            seen += idx
            sizes.append(len(idx))
        assert sorted(seen) == list(range(n))
        assert max(sizes) - min(sizes) <= 1


def test_contiguous_chunks_are_ordered_like_the_frames():
    bounds = shard.chunk_bounds(13, 4)
    assert bounds == [(0, 4), (4, 7), (7, 10), (10, 13)]
    for r in range(4):
        idx = shard.partition_contiguous(13, 4, r)
        assert idx == list(range(*bounds[r]))


def test_bad_requests_are_rejected():
    with pytest.raises(ValueError):
        shard.partition_round_robin(4, 2, 2)
    with pytest.raises(ValueError):
        shard.partition_contiguous(4, 0, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    info = shard.init_process_group("gloo")
    assert (info.rank, info.world) == (rank, world)
    jobs = [f"img{i}" for i in range(7)]
    local = shard.run_sharded(jobs, lambda i, j: (j, rank), info)
    merged = shard.run_sharded(jobs, lambda i, j: (j, rank), info, gather=True)
    frames = shard.run_sharded(list(range(9)), lambda i, j: rank, info, contiguous=True, gather=True)
    # the per-pass frame exchange of a sharded video with ragged chunks: rank 0 produced three frames, rank 1 none
    fresh = {(16, 1, f): torch.full((4, 5, 3), 10 * f + 1, dtype=torch.uint8) for f in ([0, 1, 2] if rank == 0 else [])}
    merged_frames = shard.exchange_frames(fresh, info)
    assert sorted(merged_frames) == [(16, 1, 0), (16, 1, 1), (16, 1, 2)]
    assert all(int(v[0, 0, 0]) == 10 * k[2] + 1 and v.shape == (4, 5, 3) for k, v in merged_frames.items())
    assert shard.exchange_frames({}, info) == {}  # a pass in which nobody produced anything
    shard.barrier()
    slowest = shard.max_over_ranks(1.0 + rank)
    total = shard.sum_over_ranks(len(local))
    q.put((rank, sorted(local), merged, frames, slowest, total))
    dist.destroy_process_group()


def test_world_size_2_gloo_control_plane():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    assert results[0][1] == [0, 2, 4, 6] and results[1][1] == [1, 3, 5]
    for rank, _, merged, frames, slowest, total in results:
        assert sorted(merged) == list(range(7))
        assert all(merged[i] == (f"img{i}", i % 2) for i in range(7))   # job i ran on rank i % world
        assert [frames[i] for i in range(9)] == [0] * 5 + [1] * 4       # contiguous chunks in frame order
        assert slowest == 2.0                                           # max over ranks
        assert total == 7.0


# ---- a video sharded in contiguous frame chunks (shard.stylize_video): schedule, chunk heads, the per-pass frame exchange ----
VID = dict(n=5, hw=(32, 40), sizes=[24, 40], iters=[4, 2], passes=2)


def _video_inputs():
    import numpy as np

    rs = np.random.RandomState(11)
    n, (h, w) = VID["n"], VID["hw"]
    frames = [rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for _ in range(n)]
    style = rs.randint(0, 256, size=(36, 30, 3)).astype(np.uint8)
    flows = {}
    for i in range(n):
        for d, j in (("forward", (i + 1) % n), ("backward", (i - 1) % n)):
            flows[(d, i, j)] = ((rs.randn(h // 2, w // 2, 2) * 0.8).astype(np.float32), (rs.rand(h // 2, w // 2) * 255).astype(np.uint8))
    return frames, style, flows


def _stand_in_optimize(content, pastiche, temporal):
    """A cheap deterministic stand-in for the per-frame optimisation (the arithmetic is tested elsewhere; this test is about which
    frame starts from what): numpy fp32, the same function on both sides of the comparison."""
    import numpy as np

    out = np.float32(0.5) * np.asarray(pastiche, np.float32) + np.float32(0.5) * np.asarray(content, np.float32)
    if temporal is not None:
        out = out + np.float32(0.25) * (np.asarray(temporal[0], np.float32) * np.asarray(temporal[1], np.float32) - out)
    return out.astype(np.float32)


def _video_worker(rank, world, port, q, tmpdir):
    import contextlib
    import types

    import numpy as np

    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import make_args
    from maua_style_b200 import image_ops, style
    from oracle import image_oracle as I

    info = shard.init_process_group("gloo")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    net = types.SimpleNamespace(temporal=None)
    # the device calls of the driver, replaced by the CPU oracle's (as in tests/test_host_logic.py)
    style._device = lambda args: torch.device("cpu")
    torch.cuda.device = lambda dev: contextlib.nullcontext()

    def load_model(args):
        net.temporal = None
        return net, []

    style.models.load_model = load_model
    style.optim.set_temporal_targets = lambda nt, warp, warp_weights=None, args=None: setattr(nt, "temporal", (warp.numpy(), warp_weights.numpy()))
    style.optim.optimize_device = lambda content, styles, init, iters, args, nt, losses: t(_stand_in_optimize(content.numpy(), init.numpy(), nt.temporal))
    image_ops.interpolate = lambda x, size=None, scale_factor=None: t(I.resize_bilinear(x.numpy(), size=None if size is None else tuple(size), scale_factor=scale_factor))
    image_ops.flow_warp_grid = lambda flow, size: t(I.flow_warp_map(flow.numpy(), size))[None]
    image_ops.grid_sample = lambda x, g: t(I.grid_sample_border(x.numpy()[0], g.numpy()[0]))[None]
    image_ops.blend = lambda x, y, a, b: t(I.blend(x.numpy(), y.numpy(), a, b))
    image_ops.deprocess_u8 = lambda x: t(I.deprocess_u8(x.numpy()))
    image_ops.preprocess = lambda img, device=None: t(I.preprocess_u8(img.numpy()))

    frames, style_rgb, flows = _video_inputs()
    a = make_args(Path(tmpdir) / "unused.pth", Path(tmpdir) / f"r{rank}", transfer_type="vid_img", image_sizes=VID["sizes"],
                  num_iters=VID["iters"], passes_per_scale=VID["passes"], init="prev_warp", temporal_blend=0.5, loop=False,
                  style_scale=1.0, match_histograms=False)
    # (the oracle's flow_warp_map takes the raw field and does normalisation + blur itself, so the raw field is passed through)
    get = lambda d, i, j: (t(flows[(d, i, j)][0]), t(flows[(d, i, j)][1].astype(np.float32) / np.float32(255))[None, None])
    seen = []
    store = shard.stylize_video([t(I.preprocess_u8(f)) for f in frames], [t(I.preprocess_u8(style_rgb))], a, get, info,
                                on_frame=lambda s, p, f, u8: seen.append(f))
    # the uncoupled variant (no collective): every rank styles its chunk as a clip of its own
    chunk_flows = dict(flows)
    for r in range(world):
        idx = shard.partition_contiguous(VID["n"], world, r)
        chunk_flows[("forward", idx[-1], idx[0])] = flows[("forward", idx[-1], (idx[-1] + 1) % VID["n"])]   # the pair that closes a pass
        chunk_flows[("backward", idx[0], idx[-1])] = flows[("backward", idx[0], (idx[0] - 1) % VID["n"])]
    get2 = lambda d, i, j: (t(chunk_flows[(d, i, j)][0]), t(chunk_flows[(d, i, j)][1].astype(np.float32) / np.float32(255))[None, None])
    a2 = make_args(Path(tmpdir) / "unused.pth", Path(tmpdir) / f"r{rank}", transfer_type="vid_img", image_sizes=VID["sizes"],
                   num_iters=VID["iters"], passes_per_scale=VID["passes"], init="prev_warp", temporal_blend=0.5, loop=False,
                   style_scale=1.0, match_histograms=False)
    own = shard.stylize_video_chunks([t(I.preprocess_u8(f)) for f in frames], [t(I.preprocess_u8(style_rgb))], a2, get2, info)
    q.put((rank, sorted(set(seen)), {k: v.numpy() for k, v in store.items()}, {k: v.numpy() for k, v in own.items()}))
    dist.destroy_process_group()


def test_world_size_2_gloo_sharded_video(tmp_path):
    """Two ranks, five frames in chunks [0,1,2] / [3,4]: every rank ends up with every frame of every pass, and the frames equal
    what the oracle's driver computes when the chain of carried-over results breaks at every chunk boundary."""
    import numpy as np

    from oracle import image_oracle as I

    world = 2
    for r in range(world):
        (tmp_path / f"r{r}").mkdir()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_video_worker, args=(r, world, port, q, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][1] == [0, 1, 2] and results[1][1] == [3, 4]  # who styled what
    frames, style_rgb, flows = _video_inputs()
    chunks = [shard.partition_contiguous(VID["n"], world, r) for r in range(world)]
    want = I.vid_img(frames, [I.preprocess_u8(style_rgb)], VID["sizes"], VID["iters"], VID["passes"],
                     lambda content, styles, pastiche, iters, temporal: _stand_in_optimize(content, pastiche, temporal),
                     lambda d, i, j: flows[(d, i, j)], init="prev_warp", temporal_blend=0.5, chunks=chunks)
    plain = I.vid_img(frames, [I.preprocess_u8(style_rgb)], VID["sizes"], VID["iters"], VID["passes"],
                      lambda content, styles, pastiche, iters, temporal: _stand_in_optimize(content, pastiche, temporal),
                      lambda d, i, j: flows[(d, i, j)], init="prev_warp", temporal_blend=0.5)
    assert len(want) == len(VID["sizes"]) * VID["passes"] * VID["n"]
    for rank, _, store, _own in results:
        assert sorted(store) == sorted(want)  # after the last exchange every rank holds the whole job
        for k in want:
            assert np.array_equal(store[k], want[k]), (rank, k)
    # sharding changes chunk heads only in how fresh their predecessor is: the job is not bit-identical to the unsharded one ...
    assert any(not np.array_equal(want[k], plain[k]) for k in want)
    # ... but a frame that never follows a chunk boundary in the first pass of the first scale is
    assert np.array_equal(want[(VID["sizes"][0], 1, 1)], plain[(VID["sizes"][0], 1, 1)])
    # stylize_video_chunks: every rank's result is the oracle's driver run on that chunk as a clip of its own
    for rank, _, _store, own in results:
        idx = chunks[rank]
        chunk_flows = lambda d, i, j: flows[(d, idx[i], (idx[i] + (1 if d == "forward" else -1)) % VID["n"])]
        alone = I.vid_img([frames[i] for i in idx], [I.preprocess_u8(style_rgb)], VID["sizes"], VID["iters"], VID["passes"],
                          lambda content, styles, pastiche, iters, temporal: _stand_in_optimize(content, pastiche, temporal),
                          chunk_flows, init="prev_warp", temporal_blend=0.5)
        assert sorted(own) == sorted((s, p, idx[f]) for (s, p, f) in alone)
        for (s, p, f), v in alone.items():
            assert np.array_equal(own[(s, p, idx[f])], v), (rank, s, p, f)
