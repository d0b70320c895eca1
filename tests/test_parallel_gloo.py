"""The N>1 host logic (saunet_b200.parallel) on CPU: world_size-2 gloo run of the flat gradient arena all-reduce
and the batch sharding rule.  No GPU, no CUDA library calls."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


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
    from saunet_b200.parallel import GradArena, shard_batch
    torch.manual_seed(0)
    shared = nn.Conv2d(3, 5, 3)
    m = nn.Sequential(shared, nn.BatchNorm2d(5), nn.Conv2d(5, 2, 1))
    m.alias = shared                                  # a parameter reachable under two names, like the SAUNet encoder
    m[2].bias.requires_grad_(False)                   # like encoder.classifier: never receives a gradient
    arena = GradArena(m, bucket_mb=0.0001)            # tiny buckets -> several async all-reduces
    assert len(arena.buckets) > 1
    n_unique = sum(p.numel() for p in {id(p): p for p in m.parameters() if p.requires_grad}.values())
    assert arena.flat.numel() >= n_unique
    for p in m.parameters():
        if p.requires_grad:
            assert p.grad is not None and p.grad.shape == p.shape
            assert arena.ptr(p) % 16 == 0
            p.grad.fill_(float(rank + 1))             # writes through the view into the flat buffer
    assert arena.ptr(m[2].bias) is None
    arena.all_reduce()
    expect = sum(range(1, world + 1)) / world
    ok = all(torch.allclose(p.grad, torch.full_like(p.grad, expect)) for p in m.parameters() if p.requires_grad)
    arena.zero()
    ok = ok and float(arena.flat.abs().sum()) == 0.0
    q.put((rank, ok, shard_batch(7, rank, world)))
    dist.destroy_process_group()


def test_grad_arena_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == [0, 2, 4, 6] and res[1][2] == [1, 3, 5]


def test_arena_single_process_noop():
    from saunet_b200.parallel import GradArena
    m = nn.Linear(4, 3)
    a = GradArena(m)
    m.weight.grad.fill_(2.0)
    a.all_reduce()                                    # no process group: must be a no-op
    assert float(m.weight.grad.sum()) == 24.0
