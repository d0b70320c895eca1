"""Pipelined front end of the training step (the loop body of train.py:90-118 without its per-step stall).

The reference's loop does, every iteration, `feed -> cuda`, forward, backward, optimizer step and `loss.item()`: the host
waits for the GPU at `item()` and the GPU then waits for the host to copy the next batch and to re-issue ~1100 launches.
``TrainLoop.step(host_feed)`` keeps the same per-step work -- every step's inputs are copied from (pinned) host memory
and every step's loss is read back -- but overlaps it:

  * the batch of step n is copied on a COPY stream into one of two device buffers while step n-1 still computes;
  * the launches of step n are issued before the host looks at step n-1's loss, whose device-to-host copy (into pinned
    memory) finished long before; ``step`` therefore returns the loss of the PREVIOUS step (None on the first call) and
    ``flush()`` returns the last one.

Nothing here is arithmetic: the step itself is the SegmentationModule call + backward through libsaunet_b200.so.
"""
import torch


class TrainLoop:
    def __init__(self, seg_module, arena, example_feed, optimizer=None, epoch=0):
        """seg_module: models.SegmentationModule on a CUDA device; arena: parallel.GradArena of its unet; example_feed:
        dict(image, seg, edge) with the shapes / dtypes of every later batch (host or device tensors); optimizer: a
        torch optimizer or saunet_b200.optim.FusedOptimizer (stepped after the gradient all-reduce) or None."""
        self.seg_module, self.arena, self.optimizer, self.epoch = seg_module, arena, optimizer, epoch
        self.dev = next(seg_module.parameters()).device
        self.bufs = [{k: torch.empty(v.shape, dtype=v.dtype, device=self.dev) for k, v in example_feed.items()} for _ in range(2)]
        self.copy = torch.cuda.Stream(device=self.dev)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.done = [None, None]
        self.loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        self.n = 0

    def step(self, host_feed):
        i = self.n & 1
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy):
            if self.done[i] is not None:
                self.copy.wait_event(self.done[i])          # step n-2 (the last reader of this buffer) has finished
            for k, v in host_feed.items():
                self.bufs[i][k].copy_(v, non_blocking=True)
            self.ready[i].record(self.copy)
        cur.wait_event(self.ready[i])
        b = self.bufs[i]
        self.arena.zero()
        loss, _acc = self.seg_module({"image": b["image"], "mask": (b["seg"], b["edge"])}, self.epoch)
        loss.backward()
        self.arena.all_reduce()
        if self.optimizer is not None:
            self.optimizer.step()
        self.loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.done[i] = ev
        prev = None
        if self.n > 0:
            self.done[1 - i].synchronize()
            prev = float(self.loss_host[1 - i])
        self.n += 1
        return prev

    def flush(self):
        """Wait for the last issued step and return its loss (None if no step ran)."""
        if self.n == 0:
            return None
        i = (self.n - 1) & 1
        self.done[i].synchronize()
        return float(self.loss_host[i])
