"""Data parallelism for the SAUNet hot path: one process per GPU, full parameter replica, batch axis sharded.

Replaces the reference's thread-based ``UserScatteredDataParallel`` + per-step parameter broadcast
(lib/nn/parallel/data_parallel.py:48-62, train.py:272-278) with resident replicas and ONE gradient all-reduce per
step over NCCL (NVLink 5 / NVSwitch).  Gradients live in a single flat fp32 arena: the backward kernels accumulate
straight into it (no per-parameter tensors, one memset per step), and the all-reduce runs in place on a handful of
large buckets, so its cost is launch latency + 127.5 MB over NVSwitch (~0.3 ms), not 515 small messages.

BatchNorm statistics stay per-GPU (what the reference does for 144 of its 150 norm layers under DataParallel and
for all of them on one GPU); the loss is averaged by the 1/world gradient scale, as ``loss.mean()`` over replicas
does in train.py:96.
"""
import torch
import torch.distributed as dist


class GradArena:
    """Flat fp32 gradient storage for the unique parameters of ``module``; ``p.grad`` become views into it."""

    def __init__(self, module, bucket_mb=32):
        seen, params = set(), []
        for p in module.parameters():
            if id(p) not in seen and p.requires_grad:
                seen.add(id(p))
                params.append(p)
        self.params = params
        dev = params[0].device
        self.offsets = {}
        off = 0
        for p in params:
            self.offsets[id(p)] = off
            off += (p.numel() + 3) // 4 * 4          # keep every view 16-byte aligned
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        for p in params:
            o = self.offsets[id(p)]
            p.grad = self.flat[o:o + p.numel()].view_as(p)
        n = max(1, int(bucket_mb * (1 << 20) // 4))
        self.buckets = [self.flat[i:i + n] for i in range(0, off, n)]
        module._saunet_grad_arena = self

    def ptr(self, p):
        o = self.offsets.get(id(p))
        return None if o is None else self.flat.data_ptr() + 4 * o

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, group=None):
        """SUM-reduce in place and scale by 1/world (mean over ranks)."""
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        works = [dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group, async_op=True) for b in self.buckets]
        for w in works:
            w.wait()
        self.flat.mul_(1.0 / world)


def shard_batch(n_items, rank, world):
    """Indices of the slices rank ``rank`` owns: r, r+world, r+2*world, ... (batch / z-stack axis)."""
    return list(range(rank, n_items, world))
