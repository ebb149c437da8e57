"""world_size-2 gloo test (CPU) of the multi-GPU plumbing used by bench.py: contiguous sharding by
image pair, no data-path collective, max-over-ranks timing, whole-job throughput."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pwc_net_pytorch_b200 import parallel


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 32, 33, 64):
        for world in (1, 2, 3, 4, 8):
            got = [parallel.shard_range(total, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)


def test_single_process_passthrough():
    assert parallel.max_over_ranks(3.5) == 3.5
    assert parallel.job_throughput(32, 2.0) == (16000.0, 2.0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        f1 = torch.arange(10 * 3, dtype=torch.float32).view(10, 3)
        (mine,) = parallel.shard_batch((f1,), rank, world)
        parallel.barrier()
        # rank r pretends its step took (r + 1) ms for its shard
        thr, worst = parallel.job_throughput(mine.shape[0], float(rank + 1))
        checksum = parallel.sum_over_ranks(float(mine.sum()))
        q.put((rank, mine.shape[0], thr, worst, checksum))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [5, 5]
    for _, _, thr, worst, checksum in res:
        assert worst == 2.0                      # slowest rank
        assert thr == pytest.approx(10 / 2e-3)   # all pairs / slowest time
        assert checksum == float(torch.arange(30, dtype=torch.float32).sum())   # shards partition the batch
