"""Fused multi-tensor optimizer for the SAUNet hot path.

The reference builds SGD / Adam / RAdam over two parameter groups (train.py:166-207: conv and linear weights with
weight decay, biases and BatchNorm affine parameters without) and steps them with a Python loop over ~700 tensors
(radam.py:15-78 makes fp32 copies of every tensor per step).  Here the parameters of the model live in ONE flat fp32
buffer (``GradArena.flatten_params``; each ``nn.Parameter`` keeps its identity, shape and state_dict key -- its
``.data`` becomes a view), the gradients in the arena's flat buffer, and a step is ONE launch of
``saunet_optimizer_step`` (csrc/optim.cu) over both, graph-capturable (the step counter lives on the device).

``FusedOptimizer`` is a ``torch.optim.Optimizer``: ``param_groups`` carry ``lr`` / ``weight_decay`` exactly like the
reference's optimizers, so ``adjust_learning_rate`` (train.py:210-216) and ``zero_grad`` work unchanged.
"""
import ctypes
import struct

import torch
import torch.nn as nn

from . import _C

KINDS = {"sgd": 0, "adam": 1, "radam": 2}


def group_weight(module):
    """train.py:166-185: [decay group (conv / linear weights), no-decay group (biases, BatchNorm weight + bias)]."""
    decay, no_decay = [], []
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.modules.conv._ConvNd)):
            decay.append(m.weight)
            if m.bias is not None:
                no_decay.append(m.bias)
        elif isinstance(m, nn.modules.batchnorm._BatchNorm):
            if m.weight is not None:
                no_decay.append(m.weight)
            if m.bias is not None:
                no_decay.append(m.bias)
    return [dict(params=decay), dict(params=no_decay, weight_decay=0.0)]


class FusedOptimizer(torch.optim.Optimizer):
    """kind: 'sgd' (momentum, nesterov=False), 'adam' or 'radam' (the reference's radam.RAdam)."""

    def __init__(self, params, arena, kind="sgd", lr=1e-3, momentum=0.0, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if kind not in KINDS:
            raise ValueError("optimizer must be one of %s" % sorted(KINDS))
        super().__init__(params, dict(lr=lr, momentum=momentum, betas=betas, eps=eps, weight_decay=weight_decay))
        self.kind, self.arena = kind, arena
        arena.flatten_params()
        dev = arena.flat.device
        n = arena.n_grad
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev) if (kind != "sgd" or momentum != 0.0) else None
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev) if kind != "sgd" else None
        self.step_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self._table_key, self._table = None, None
        seen = set()
        for g in self.param_groups:
            for p in g["params"]:
                if id(p) not in arena.offsets:
                    raise ValueError("FusedOptimizer: a parameter is not part of the gradient arena")
                seen.add(id(p))
        # parameters of the arena that are in no group (never updated: lr 0) are legal, e.g. frozen layers

    def _segments(self):
        """Device table {begin, lr, weight_decay} per arena tensor, rebuilt only when a group's lr / weight decay moved."""
        key = tuple((g["lr"], g["weight_decay"]) for g in self.param_groups)
        if key == self._table_key:
            return self._table
        hp = {}
        for g in self.param_groups:
            for p in g["params"]:
                hp[id(p)] = (float(g["lr"]), float(g["weight_decay"]))
        raw = bytearray()
        for p in self.arena.params:                     # arena order = ascending offsets
            lr, wd = hp.get(id(p), (0.0, 0.0))
            raw += struct.pack("<qff", self.arena.offsets[id(p)], lr, wd)
        t = torch.frombuffer(raw, dtype=torch.uint8).to(self.arena.flat.device, non_blocking=False)
        if self._table is not None and self._table.numel() == t.numel():
            self._table.copy_(t)                        # in place: a captured graph reads this address
        else:
            self._table = t
        self._table_key = key
        return self._table

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        a = self.arena
        g0 = self.param_groups[0]
        tab = self._segments()
        _C.call("saunet_optimizer_step", KINDS[self.kind], a.params_flat.data_ptr(), a.flat.data_ptr(),
                self.exp_avg.data_ptr() if self.exp_avg is not None else None,
                self.exp_avg_sq.data_ptr() if self.exp_avg_sq is not None else None, a.n_grad, tab.data_ptr(),
                len(a.params), self.step_counter.data_ptr(), float(g0["betas"][0]), float(g0["betas"][1]),
                float(g0["eps"]), float(g0["momentum"]), torch.cuda.current_stream(a.flat.device).cuda_stream,
                nbytes=28.0 * a.n_grad)
        return loss

    def zero_grad(self, set_to_none=False):
        """One memset of the flat arena (the views stay attached)."""
        self.arena.zero()


def create_fused_optimizer(unet, arena, name, lr, momentum=0.9, weight_decay=1e-4, betas=(0.9, 0.999)):
    """train.py:188-207 (create_optimizers) on the fused path."""
    name = name.lower()
    groups = group_weight(unet)
    in_arena = set(arena.offsets)
    for g in groups:
        g["params"] = [p for p in g["params"] if id(p) in in_arena]
    if name == "sgd":
        return FusedOptimizer(groups, arena, "sgd", lr=lr, momentum=momentum, weight_decay=weight_decay)
    if name == "adam":
        return FusedOptimizer(groups, arena, "adam", lr=lr, betas=betas)
    if name == "radam":
        return FusedOptimizer(groups, arena, "radam", lr=lr, betas=betas)
    raise ValueError("unknown optimizer %r" % name)
