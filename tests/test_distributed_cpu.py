"""Multi-GPU host logic on CPU: world_size-2 gloo run of the shard arithmetic and the
per-rollout episode-statistics reduction (the only collective in the design)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge

    pcbgpu = ge.build()
    lo, hi = pcbgpu.shard_range(1000, rank, world)
    local = {"episodes": float(hi - lo), "sum_return": -1.5 * (rank + 1), "sum_length": 10.0 * (hi - lo),
             "sum_wirelength": 0.25 * rank, "sum_intersections": 2.0, "steps": 100.0}
    total = pcbgpu.allreduce_stats(local)
    out[rank] = (lo, hi, total)
    dist.destroy_process_group()


def test_shards_and_stat_reduction_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (lo0, hi0, t0), (lo1, hi1, t1) = out[0], out[1]
    assert (lo0, hi0, lo1, hi1) == (0, 500, 500, 1000)
    assert t0 == t1
    assert t0["episodes"] == 1000.0 and t0["sum_return"] == -4.5 and t0["steps"] == 200.0
    assert t0["sum_wirelength"] == 0.25


def test_shard_range_covers_everything():
    import __graft_entry__ as ge

    pcbgpu = ge.build()
    for total, world in [(10, 3), (65536 * 8, 8), (7, 8)]:
        spans = [pcbgpu.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
