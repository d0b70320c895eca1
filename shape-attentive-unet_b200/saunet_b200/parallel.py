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
        # second half of the allocation: the conv kernels' PACKED weight-gradient images ([(tap, cin)][cout], what the
        # tensor-core wgrad epilogues write with coalesced atomics); one table-driven kernel folds all of them into the
        # parameter-layout gradients at the end of backward (instead of a zero-fill + an unpack launch per conv)
        self.n_grad = off
        self.packed_off = {}
        entries, first = [], 0
        for p in params:
            if p.dim() == 4:
                self.packed_off[id(p)] = off
                a, b, kh, kw = p.shape
                entries.append((first, off - self.n_grad, self.offsets[id(p)], a, b, kh * kw))
                first += p.numel()
                off += (p.numel() + 3) // 4 * 4
        self.flat_all = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat = self.flat_all[:self.n_grad]
        self.packed = self.flat_all[self.n_grad:]
        self._unpack_total = first
        self._unpack_n = len(entries)
        if entries:
            import ctypes
            from . import _C
            arr = (_C.UnpackEntry * len(entries))()
            for i, e in enumerate(entries):
                arr[i].first, arr[i].packed_off, arr[i].grad_off, arr[i].A, arr[i].Bc, arr[i].T = e
            raw = bytes(arr)
            self._unpack_table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self._views = {}
        for p in params:
            o = self.offsets[id(p)]
            self._views[id(p)] = self.flat[o:o + p.numel()].view_as(p)
            p.grad = self._views[id(p)]
        # Buckets = contiguous parameter ranges of ~bucket_mb each.  Parameters are laid out in forward (module) order, so
        # backward completes them from the END of the arena: the last bucket is final first.  Each bucket carries the
        # unpack table of its own conv weights, so that "fold the packed weight gradients + all-reduce" can be issued
        # per bucket on a side stream while the rest of backward still runs (see bucket_ready / Tape.backward).
        n = max(1, int(bucket_mb * (1 << 20) // 4))
        self.bucket_params, cur, start = [], [], 0
        for p in params:
            cur.append(p)
            end = self.offsets[id(p)] + (p.numel() + 3) // 4 * 4
            if end - start >= n:
                self.bucket_params.append((start, end, cur))
                cur, start = [], end
        if cur:
            self.bucket_params.append((start, self.n_grad, cur))
        self.buckets = [self.flat[a:b] for a, b, _ in self.bucket_params]
        self._bucket_of = {id(p): i for i, (_, _, ps) in enumerate(self.bucket_params) for p in ps}
        self._bucket_tables = []
        for _, _, ps in self.bucket_params:
            ents, first = [], 0
            for p in ps:
                if p.dim() == 4:
                    a_, b_, kh, kw = p.shape
                    ents.append((first, self.packed_off[id(p)] - self.n_grad, self.offsets[id(p)], a_, b_, kh * kw))
                    first += p.numel()
            self._bucket_tables.append((self._make_table(ents, dev) if ents else None, len(ents), first))
        self._schedule = {}            # n_ops -> {op index -> [bucket, ...]}, learnt from the first backward of each tape length
        self._comm = None              # side stream for unpack + all-reduce
        self._works = []
        self._overlapped = False
        self.comm_enabled = True       # False (see no_sync): bucket_ready only folds the packed gradients, no collective
        module._saunet_grad_arena = self

    @staticmethod
    def _make_table(entries, dev):
        from . import _C
        arr = (_C.UnpackEntry * len(entries))()
        for i, e in enumerate(entries):
            arr[i].first, arr[i].packed_off, arr[i].grad_off, arr[i].A, arr[i].Bc, arr[i].T = e
        return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)

    # ---- overlap of gradient post-processing with backward -----------------------------------------------------
    def learn_schedule(self, n_ops, last_touch):
        """last_touch: {id(param): index (in execution order) of the last backward op that wrote its gradient}.
        A bucket is final after the latest such op among its parameters (-1: never written, e.g. unused heads)."""
        sched = {}
        for b, (_, _, ps) in enumerate(self.bucket_params):
            k = max([last_touch.get(id(p), -1) for p in ps] + [-1])
            sched.setdefault(max(k, 0), []).append(b)
        # one schedule PER tape length: a differently shaped pass (e.g. the serialised profiling pass only rank 0 runs) must
        # not evict the schedule of the regular step -- the ranks would then disagree on whether a backward overlaps its
        # collectives, i.e. on the ORDER of the collectives
        self._schedule[n_ops] = sched

    def schedule_for(self, n_ops):
        return self._schedule.get(n_ops)

    def _world(self, group=None):
        return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1

    def bucket_ready(self, b):
        """Called from Tape.backward right after the op that finalises bucket ``b``: on the side stream, fold the bucket's
        packed conv-weight gradients into parameter layout and (N > 1, not under graph capture) start its all-reduce --
        both overlap the remaining backward kernels."""
        dev = self.flat.device
        cur = torch.cuda.current_stream(dev)
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._comm.wait_event(ev)
        capturing = torch.cuda.is_current_stream_capturing()
        with torch.cuda.stream(self._comm):
            tab, n, total = self._bucket_tables[b]
            if n:
                from . import _C
                _C.call("saunet_unpack_wgrad_multi", tab.data_ptr(), n, total, self.packed.data_ptr(), self.flat.data_ptr(),
                        self._comm.cuda_stream)
            world = self._world()
            if world > 1 and not capturing and self.comm_enabled:
                op = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
                self._works.append((dist.all_reduce(self.buckets[b], op=op, async_op=True), b, op))
        self._overlapped = True

    def no_sync(self):
        """Context manager: backward passes inside it issue NO collective (like DistributedDataParallel.no_sync) -- for
        gradient accumulation, and for anything only some ranks execute (e.g. a rank-0-only profiling pass: a
        collective the other ranks never enter would deadlock the job)."""
        arena = self

        class _NoSync:
            def __enter__(self_):
                self_.prev, arena.comm_enabled = arena.comm_enabled, False

            def __exit__(self_, *a):
                arena.comm_enabled = self_.prev
        return _NoSync()

    def join(self):
        """Make the current stream wait for everything issued by bucket_ready (end of backward / before the optimizer)."""
        if self._comm is None:
            return
        dev = self.flat.device
        world = self._world()
        with torch.cuda.stream(self._comm):
            for w, b, op in self._works:
                w.wait()
                if op == dist.ReduceOp.SUM:
                    self.buckets[b].mul_(1.0 / world)
        torch.cuda.current_stream(dev).wait_stream(self._comm)

    def flatten_params(self):
        """Move every parameter of the arena into ONE flat fp32 buffer laid out like the gradient arena (same
        offsets), re-pointing ``p.data`` at views of it: ``nn.Parameter`` identities, shapes and state_dict keys are
        unchanged (load_state_dict copies into the views).  What the fused optimizer (saunet_b200.optim) steps over."""
        if getattr(self, "params_flat", None) is not None:
            return self.params_flat
        flat = torch.zeros(self.n_grad, dtype=torch.float32, device=self.flat.device)
        with torch.no_grad():
            for p in self.params:
                o = self.offsets[id(p)]
                v = flat[o:o + p.numel()].view_as(p)
                v.copy_(p.data)
                p.data = v
        self.params_flat = flat
        from . import engine
        engine.invalidate_packed()          # data pointers moved: every cached packed image is keyed on them
        return flat

    def ptr(self, p):
        o = self.offsets.get(id(p))
        return None if o is None else self.flat.data_ptr() + 4 * o

    def packed_ptr(self, p):
        """Device pointer of the packed weight-gradient image of conv weight ``p`` (None: not a conv weight here)."""
        o = self.packed_off.get(id(p))
        return None if o is None else self.flat_all.data_ptr() + 4 * o

    def unpack(self, stream):
        """Fold every packed conv weight gradient into its parameter-layout gradient (one launch)."""
        if self._unpack_n:
            from . import _C
            _C.call("saunet_unpack_wgrad_multi", self._unpack_table.data_ptr(), self._unpack_n, self._unpack_total,
                    self.packed.data_ptr(), self.flat.data_ptr(), stream)

    def zero(self):
        self.flat.zero_()          # (the packed images are cleared by the unpack kernel as it consumes them)
        self.attach()

    def attach(self):
        """(Re-)install the arena views as ``p.grad``."""
        for p in self.params:
            v = self._views[id(p)]
            if p.grad is not v:
                p.grad = v

    def ensure_attached(self):
        """Called at the start of every backward.  ``module.zero_grad()`` / ``optimizer.zero_grad()`` default to
        set_to_none=True (train.py:93 of the reference calls the former every iteration): ``p.grad`` is then None while
        the kernels keep accumulating into the arena, and optimizer.step() would silently skip every parameter.  A
        detached view therefore means "the gradients were just reset": zero the arena and re-attach the views."""
        if any(p.grad is not self._views[id(p)] for p in self.params):
            self.flat.zero_()
            self.attach()

    def all_reduce(self, group=None):
        """Mean of the gradients over ranks, in place.  When the backward that just ran overlapped its buckets
        (bucket_ready), the all-reduces are already in flight: only wait for them."""
        if not (dist.is_available() and dist.is_initialized()):
            self._works = []
            return
        world = dist.get_world_size(group)
        if world == 1:
            self._works = []
            return
        if self._works:
            self.join()
            self._works = []
            return
        nccl = dist.get_backend(group) == "nccl"
        op = dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM
        works = [dist.all_reduce(b, op=op, group=group, async_op=True) for b in reversed(self.buckets)]
        for w in works:
            w.wait()
        if not nccl:
            self.flat.mul_(1.0 / world)


def shard_batch(n_items, rank, world):
    """Indices of the slices rank ``rank`` owns: r, r+world, r+2*world, ... (batch / z-stack axis)."""
    return list(range(rank, n_items, world))
