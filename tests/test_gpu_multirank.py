"""N-rank data parallelism on the GPU: two ranks, each running half of a batch through the CUDA
path with the flat GradArena, bucketed unpack + all-reduce overlapped with backward (saunet_b200.parallel), must end up
with the gradients a single process computes on the concatenated batch (BatchNorm in eval mode, so that batch
statistics do not couple the samples; a loss that is a plain mean, so that the mean of per-rank losses IS the global
loss).  Two processes share cuda:0 over gloo when the box has one GPU (NCCL refuses two ranks on one device); with two
or more GPUs the same test runs over NCCL, one rank per GPU."""
import os
import socket
import warnings

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev):
    from helpers import template_state_dict
    from models import SAUNet
    from saunet_b200 import synth
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = SAUNet(num_classes=4, pretrained=False)
    m.load_state_dict(synth.synthetic_state_dict(template_state_dict(), seed=0))
    return m.to(dev).eval()            # eval-mode BatchNorm; parameters still require grad


def _loss(seg, edge, cs, ce):
    return (seg * cs).mean() + (edge * ce).mean()


def _worker(rank, world, port, backend, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here), os.path.join(os.path.dirname(here), "shape-attentive-unet_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    from saunet_b200 import synth
    from saunet_b200.parallel import GradArena
    data = synth.synthetic_batch(4, 64, seed=304)
    g = torch.Generator().manual_seed(1)
    cs, ce = torch.randn(4, 4, 64, 64, generator=g), torch.randn(4, 1, 64, 64, generator=g)
    m = _build(dev)
    arena = GradArena(m, bucket_mb=8)
    assert len(arena.buckets) > 4
    mine = slice(rank * 2, rank * 2 + 2)
    out = {}
    for step in range(3):              # step 0 learns the bucket schedule, steps 1-2 overlap unpack + all-reduce with backward
        arena.zero()
        seg, edge = m(data["image"][mine].to(dev))
        _loss(seg, edge, cs[mine].to(dev), ce[mine].to(dev)).backward()
        overlapped = bool(arena._works)
        arena.all_reduce()
        torch.cuda.synchronize()
        out[step] = (overlapped, arena.flat.clone())
    res = {"rank": rank, "overlapped": [out[s][0] for s in range(3)]}
    res["steps_agree"] = float((out[2][1] - out[0][1]).norm() / out[0][1].norm())
    if rank == 0:
        ref = _build(dev)              # single process, whole batch, plain autograd-returned gradients (no arena)
        seg, edge = ref(data["image"].to(dev))
        _loss(seg, edge, cs.to(dev), ce.to(dev)).backward()
        torch.cuda.synchronize()
        num = den = 0.0
        worst = 0.0
        for (k, p), pr in zip(m.named_parameters(), ref.parameters()):
            if pr.grad is None:
                continue
            d = (p.grad - pr.grad).double()
            num += float(d.pow(2).sum()); den += float(pr.grad.double().pow(2).sum())
            if k in ("final.weight", "dec0.0.weight", "encoder.features.conv0.weight", "gate1.weight", "dec3.c3x3rb.0.weight"):
                worst = max(worst, float(d.norm() / pr.grad.double().norm().clamp_min(1e-30)))
        res["rel"] = (num / den) ** 0.5
        res["worst_named"] = worst
    q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_equal_single_process():
    world = 2
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0 = [r for r in res if r["rank"] == 0][0]
    for r in res:
        assert r["overlapped"] == [False, True, True], r       # first step learns, later steps overlap
        assert r["steps_agree"] < 2e-3, r                      # same gradients with and without the overlap
    # eval-mode BN + linear loss: the averaged 2-rank gradients ARE the single-process gradients (up to fp32 atomics)
    assert r0["rel"] < 2e-3 and r0["worst_named"] < 2e-3, r0
