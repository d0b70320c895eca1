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


def _vol_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from saunet_b200 import inference as inf
    # the z-stack sharding + gather of predict_volume with a stand-in for the CUDA forward: label = slice index
    Z, H, W = 7, 4, 5
    mine = inf.shard_slices(Z, rank, world)
    lab = torch.stack([torch.full((H, W), z, dtype=torch.uint8) for z in mine])
    longest = (Z + world - 1) // world
    pad = torch.zeros((longest, H, W), dtype=torch.uint8)
    pad[:lab.shape[0]] = lab
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    vol = inf.merge_gathered(parts, Z, world)
    ok = all(int(vol[z].min()) == z == int(vol[z].max()) for z in range(Z))
    q.put((rank, ok, mine))
    dist.destroy_process_group()


def test_volume_slices_shard_and_gather_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_vol_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == [0, 2, 4, 6] and res[1][2] == [1, 3, 5]


def test_packing_geometry():
    """undo_crop / resample_nearest (test_and_pack.py:31-76): a centre crop is undone by zero padding, a zero padding by
    centre cropping, and the order-0 resample is the identity at equal shapes and picks nearest voxels otherwise."""
    import numpy as np
    from saunet_b200 import inference as inf
    pred = np.arange(1, 1 + 6 * 8, dtype=np.uint8).reshape(6, 8)
    big = inf.undo_crop((10, 13), pred)              # original slice was larger than the 6x8 network input: it was cropped
    assert big.shape == (10, 13) and big.sum() == pred.sum()
    y, x = np.argwhere(big == 1)[0]
    assert (y, x) == (2, 3) and np.array_equal(big[y:y + 6, x:x + 8], pred)     # centred (round-half-up offsets)
    small = inf.undo_crop((4, 5), pred)              # original was smaller: it was zero padded, so crop the centre back out
    assert small.shape == (4, 5) and np.array_equal(small, pred[1:5, 1:6])
    vol = np.random.default_rng(0).integers(0, 4, (6, 8, 3)).astype(np.uint8)
    assert np.array_equal(inf.resample_nearest(vol, vol.shape), vol)
    up = inf.resample_nearest(vol, (12, 16, 3))
    assert up.shape == (12, 16, 3) and np.array_equal(up[::2, ::2], vol) and set(np.unique(up)) <= {0, 1, 2, 3}
