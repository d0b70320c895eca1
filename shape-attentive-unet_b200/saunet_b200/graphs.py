"""CUDA-graph capture of one whole training step (forward + DualLoss + metrics + backward).

The step is ~1300 short kernel launches issued from Python; eagerly the GPU idles at every host synchronisation
(`loss.item()` in train.py:114) while the host re-issues them.  ``GraphedStep`` captures the launches once and
replays them with a single `cudaGraphLaunch`: inputs are copied into static device buffers (the H2D copy of the
step), gradients land in the module's flat ``GradArena`` and the loss in a static tensor.

Only launches are captured -- the arithmetic is the same libsaunet_b200.so kernels.  The weight-packing cache is
bypassed during capture so the pack kernels are part of the graph (weights change every optimizer step).
"""
import torch

from . import engine


class GraphedStep:
    def __init__(self, seg_module, arena, example_feed, epoch=0, warmup=2, optimizer=None):
        """seg_module: models.SegmentationModule on a CUDA device; arena: parallel.GradArena of its unet;
        example_feed: dict(image, seg, edge) device tensors with the shapes of every later step;
        optimizer: a saunet_b200.optim.FusedOptimizer to make the weight update part of the captured step (its step
        counter and hyper-parameter table live on the device; learning-rate changes reach the graph through that
        table).  The warm-up steps run forward + backward only, so they do not move the weights."""
        self.seg_module, self.arena, self.optimizer = seg_module, arena, optimizer
        self.static = {k: v.clone() for k, v in example_feed.items()}
        dev = self.static["image"].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        # (no collectives in the warm-up or the capture: the gradient all-reduce follows each replay -- arena.all_reduce();
        #  an async all-reduce left in flight here would be joined from inside the capture and invalidate it)
        with torch.cuda.stream(side), arena.no_sync():    # warm-up on the side stream (lazy inits, allocator pools)
            for _ in range(warmup):
                self._eager(epoch, False)
        if optimizer is not None:
            optimizer._segments()                          # build the device table outside the capture (pageable H2D copy)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        engine.FORCE_PACK = True
        try:
            with arena.no_sync(), torch.cuda.graph(self.graph):
                self.loss, self.acc = self._eager(epoch)
        finally:
            engine.FORCE_PACK = False
            engine.reset_capture_flags()

    def _eager(self, epoch, with_optimizer=True):
        s = self.static
        self.arena.zero()
        loss, acc = self.seg_module({"image": s["image"], "mask": (s["seg"], s["edge"])}, epoch)
        loss.backward()
        if self.optimizer is not None and with_optimizer:
            self.optimizer.step()
        return loss.detach(), acc

    def __call__(self, feed):
        """feed tensors may live on the host (pinned) or the device; returns the (static) loss tensor."""
        if self.optimizer is not None:
            self.optimizer._segments()                     # refresh lr / weight decay in place if a param_group changed
        for k, v in feed.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.loss
