"""Host-side sharding logic (maua_style_b200/shard.py): partitions, and the world_size-2 control plane over gloo.

No GPU and no compute calls here: the jobs are stand-ins; the data path of the real thing has no collective at all.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maua_style_b200 import shard


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
