"""Host-side execution engine for the SAUNet hot path.

Activations live in NHWC fp32 device buffers (``Buf`` = a channel slice of a
``Store``); every arithmetic step is one call into libsaunet_b200.so through
``_C.call``.  A ``Tape`` records one backward closure per forward step, so a
whole model (or a single block) runs as ONE ``torch.autograd.Function`` whose
backward replays the tape in reverse: fwd+bwd are hand-written per block, torch
autograd only sees the module boundary.

torch is used here for device memory (caching allocator), the current stream
and the autograd boundary -- never for arithmetic on the path.  There is no CPU
fallback: a non-CUDA tensor or a missing library raises.
"""
import contextlib
import ctypes
import os
import weakref

import torch

from . import _C
from ._C import ACT_NONE, ACT_RELU, ACT_SIGMOID, ConvDesc, WgradDesc

BN_EPS = 1e-5
OVERLAP = os.environ.get("SAUNET_OVERLAP", "1") == "1"      # post-process gradient buckets on a side stream during backward
# Small problems (DenseNet blocks 3-4, the deep decoder stages: <= 128 tiles on 148 SMs, 15-30 us per launch whatever the
# size) leave most of the GPU idle; independent launches -- the weight gradient and the data gradient of one conv -- are
# therefore issued on different streams so that they share it.  SAUNET_CONCURRENCY=0 serialises everything again.
CONCURRENCY = os.environ.get("SAUNET_CONCURRENCY", "1") == "1"
ASYNC_WGRAD_MAX_PIX = int(os.environ.get("SAUNET_ASYNC_WGRAD_MAX_PIX", str(1 << 40)))     # (measured: every size gains; 50.5 -> 46.7 ms/step)
_SIDE_STREAMS = {}


def _round4(n):
    return (n + 3) // 4 * 4


class Store:
    """One NHWC allocation [npix, ld] plus (lazily) its gradient twin."""
    __slots__ = ("t", "g", "ld", "npix", "ab", "tag")

    def __init__(self, t, npix, ld, tag=0):
        self.t, self.g, self.ld, self.npix, self.tag = t, None, ld, npix, tag
        self.ab = 0          # device pointer of double[2][ld]: pending BatchNorm-backward mean terms (see bn_dgrad_fused)


class Buf:
    """Channels [c0, c0+C) of a Store, viewed as a [B,H,W,C] feature map."""
    __slots__ = ("s", "c0", "C", "B", "H", "W")

    def __init__(self, s, c0, C, B, H, W):
        self.s, self.c0, self.C, self.B, self.H, self.W = s, c0, C, B, H, W

    @property
    def ptr(self):
        return self.s.t.data_ptr() + 4 * self.c0

    @property
    def ld(self):
        return self.s.ld

    @property
    def npix(self):
        return self.B * self.H * self.W

    def slice(self, c0, C):
        assert 0 <= c0 and c0 + C <= self.C
        return Buf(self.s, self.c0 + c0, C, self.B, self.H, self.W)

    def full(self):
        return self.c0 == 0 and self.C == self.s.ld

    def nchw(self):
        """Zero-copy NCHW-shaped (channels_last-strided) tensor view."""
        v = self.s.t.view(self.B, self.H, self.W, self.s.ld)
        if not self.full():
            v = v[..., self.c0:self.c0 + self.C]
        return v.permute(0, 3, 1, 2)


class Tape:
    def __init__(self, device, record):
        if device.type != "cuda":
            raise RuntimeError("saunet_b200 runs on CUDA devices only (got %s); there is no CPU fallback" % device)
        _C.load()
        self.device = device
        self.record = record
        self.stream = torch.cuda.current_stream(device).cuda_stream
        self.ops = []
        self.pgrads = {}
        self._dchunks = {}
        self._fchunk = None
        self._foff = 0
        self._keep = []
        self.bn_tracked = []
        self.arena = None
        self.used_packed = False
        self._touch, self._cur = None, 0
        self._dirty, self._hold, self._wg_rr = set(), [], 0
        self._tag, self._open_sections = 0, []          # forward: tag new Stores / ops are registered with; sections awaiting their join
        self._exec_tag, self._fork_ev, self._priv = 0, {}, {}     # backward: tag of the running op, fork events, private grad buffers
        self.repacked = []      # pack-cache entries refreshed under FORCE_PACK by this tape (flag reset at the end)

    # ---- streams --------------------------------------------------------
    @contextlib.contextmanager
    def side_section(self, k):
        """A forward SECTION on side stream ``k`` whose backward ops also run on that stream, concurrently with the
        backward of whatever the main stream ran between the end of this section and the next ``join_sides()``:
        the reverse of the forward join is the backward fork (an event on the main stream the side stream waits for),
        the reverse of the forward fork is the backward join.  Gradients the section's backward accumulates into
        tensors created OUTSIDE it (shared encoder features) go to private zero-initialised buffers that the backward
        join adds into the real gradients, so the two streams never read-modify-write the same memory."""
        if not CONCURRENCY:
            yield
            return
        if self.record:
            self.ops.append(("join", k))
        prev_tag, self._tag = self._tag, k
        self._open_sections.append(k)
        try:
            with self.on_side(k):
                yield
        finally:
            self._tag = prev_tag

    @contextlib.contextmanager
    def on_side(self, k, hold=()):
        """Issue the enclosed launches on side stream ``k``, ordered after everything issued so far on the current
        stream.  ``hold``: objects (Stores) the side work reads that must stay alive until the next ``join_sides``."""
        if not CONCURRENCY:
            yield
            return
        cur = torch.cuda.current_stream(self.device)
        key = (self.device.index, k)
        s = _SIDE_STREAMS.get(key)
        if s is None:
            s = _SIDE_STREAMS[key] = torch.cuda.Stream(device=self.device)
        s.wait_stream(cur)
        self._hold.extend(hold)
        prev = self.stream
        with torch.cuda.stream(s):
            self.stream = s.cuda_stream
            try:
                yield
            finally:
                self.stream = prev
        self._dirty.add(k)

    def join_sides(self):
        """The current stream waits for all side-stream work issued since the last join."""
        if self._open_sections and self.record and self._tag == 0:
            for k in self._open_sections:
                self.ops.append(("fork", k))
        if self._tag == 0:
            self._open_sections = []
        if self._dirty:
            cur = torch.cuda.current_stream(self.device)
            for k in self._dirty:
                cur.wait_stream(_SIDE_STREAMS[(self.device.index, k)])
            self._dirty.clear()
        self._hold.clear()

    # ---- memory ---------------------------------------------------------
    def new(self, B, H, W, C, ld=None):
        ld = C if ld is None else ld
        t = torch.empty(B * H * W * ld, dtype=torch.float32, device=self.device)
        return Buf(Store(t, B * H * W, ld, self._tag), 0, C, B, H, W)

    def wrap(self, t, B, H, W, C):
        """Adopt an existing contiguous [B,H,W,C] fp32 tensor."""
        return Buf(Store(t, B * H * W, C), 0, C, B, H, W)

    def dzeros(self, n):
        """n zeroed doubles -> device pointer (arena; one memset per chunk).  Chunks are per stream: the memset that
        zeroes a chunk is ordered only with the stream it was issued on."""
        n = (n + 1) // 2 * 2
        ent = self._dchunks.get(self.stream)
        if ent is None or ent[1] + n > ent[0].numel():
            t = torch.zeros(max(1 << 15, n), dtype=torch.float64, device=self.device)
            self._keep.append(t)
            ent = self._dchunks[self.stream] = [t, 0]
        p = ent[0].data_ptr() + 8 * ent[1]
        ent[1] += n
        return p

    def fempty(self, n):
        n = _round4(n)
        if self._fchunk is None or self._foff + n > self._fchunk.numel():
            self._fchunk = torch.empty(max(1 << 16, n), dtype=torch.float32, device=self.device)
            self._keep.append(self._fchunk)
            self._foff = 0
        p = self._fchunk.data_ptr() + 4 * self._foff
        self._foff += n
        return p

    def scratch(self, n):
        t = torch.empty(n, dtype=torch.float32, device=self.device)
        return t

    # ---- gradients --------------------------------------------------------
    def grad(self, buf):
        """Existing gradient of ``buf`` (a Buf over the grad Store) or None."""
        g = buf.s.g
        if g is None:
            return None
        return Buf(g, buf.c0, buf.C, buf.B, buf.H, buf.W)

    def gw(self, buf):
        """Gradient write target for ``buf`` -> (gbuf, accumulate_flag).
        First write covering the whole Store skips the zero fill."""
        s = buf.s
        if self._exec_tag and s.tag != self._exec_tag:
            # an op of a side section accumulating into a tensor it shares with the main stream: private buffer
            key = (id(s), buf.c0, buf.C)
            ent = self._priv.get(key)
            if ent is None:
                t = torch.zeros(buf.npix * buf.C, dtype=torch.float32, device=self.device)
                self._keep.append(t)
                ent = self._priv[key] = (Buf(Store(t, buf.npix, buf.C, self._exec_tag), 0, buf.C, buf.B, buf.H, buf.W), buf, self._exec_tag)
            return ent[0], 1
        if s.g is None:
            if buf.full():
                s.g = Store(torch.empty(s.npix * s.ld, dtype=torch.float32, device=self.device), s.npix, s.ld)
                return Buf(s.g, buf.c0, buf.C, buf.B, buf.H, buf.W), 0
            s.g = Store(torch.zeros(s.npix * s.ld, dtype=torch.float32, device=self.device), s.npix, s.ld)
        return Buf(s.g, buf.c0, buf.C, buf.B, buf.H, buf.W), 1

    def seed(self, buf, gt):
        """Install an incoming NHWC-contiguous gradient tensor as buf's grad."""
        s = buf.s
        assert buf.full() and s.g is None
        s.g = Store(gt, s.npix, s.ld)

    def pgrad(self, p):
        """Device pointer gradients of parameter ``p`` are accumulated into: a view of the module's flat
        GradArena when one is attached (saunet_b200.parallel), else a per-parameter zero tensor handed to autograd."""
        if self.arena is not None:
            ptr = self.arena.ptr(p)
            if ptr is not None:
                if self._touch is not None:
                    self._touch[id(p)] = self._cur
                return ptr
        g = self.pgrads.get(p)
        if g is None:
            g = torch.zeros_like(p, memory_format=torch.contiguous_format)
            self.pgrads[p] = g
        return g.data_ptr()

    def packed_grad(self, p):
        """Packed weight-gradient target of conv weight ``p`` inside the attached GradArena (zeroed with the arena,
        unpacked once at the end of backward), or None -> per-conv temporary + unpack launch."""
        if self.arena is not None:
            ptr = self.arena.packed_ptr(p)
            if ptr is not None:
                self.used_packed = True
                if self._touch is not None:
                    self._touch[id(p)] = self._cur
                return ptr
        return None

    def backward(self):
        """Replay the tape in reverse.  With a GradArena attached, the arena's buckets are post-processed (packed
        conv-weight gradients folded, gradients all-reduced when N > 1) on a side stream as soon as the LAST op that
        writes into them has run, overlapping the rest of backward; the op index per bucket is learnt from the first
        backward of a tape of this length (the op sequence of a model is static)."""
        arena = self.arena
        n = len(self.ops)
        sched = arena.schedule_for(n) if (arena is not None and OVERLAP) else None
        self._touch = {} if (arena is not None and sched is None) else None
        for i, fn in enumerate(reversed(self.ops)):
            self._cur = i
            if isinstance(fn, tuple):
                if isinstance(fn[0], str):
                    self._run_marker(*fn)
                else:
                    self._run_tagged(*fn)
            else:
                fn()
            if sched is not None and i in sched:
                self.join_sides()                  # weight gradients issued on side streams belong to the bucket too
                for b in sched[i]:
                    arena.bucket_ready(b)
        self.join_sides()
        self.ops = []
        if arena is not None:
            if sched is None:
                if self.used_packed:
                    arena.unpack(self.stream)
                if self._touch is not None and n:
                    arena.learn_schedule(n, self._touch)
            elif not arena._works or torch.cuda.is_current_stream_capturing():
                arena.join()                       # nothing left in flight but the unpack kernels: join now
        self.used_packed = False

    def on_backward(self, fn):
        if self.record:
            self.ops.append((fn, self._tag) if self._tag else fn)

    def _run_marker(self, kind, k):
        """Backward of the forward fork / join of side section ``k`` (see side_section)."""
        cur = torch.cuda.current_stream(self.device)
        if kind == "fork":                       # reverse of the forward join: the side stream may start from here
            ev = torch.cuda.Event()
            ev.record(cur)
            self._fork_ev[k] = ev
            return
        # "join": reverse of the forward fork -- wait for the section's backward, then fold its private gradient buffers
        side = _SIDE_STREAMS.get((self.device.index, k))
        if side is not None:
            cur.wait_stream(side)
        self._dirty.discard(k)
        for key in [key for key, ent in self._priv.items() if ent[2] == k]:
            pb, orig, _ = self._priv.pop(key)
            dst, acc = self.gw(orig)
            copy_slice(self, pb, dst, acc)

    def _run_tagged(self, fn, k):
        """Run a backward op of side section ``k`` on its stream, ordered after the section's fork event only."""
        side = _SIDE_STREAMS.get((self.device.index, k))
        ev = self._fork_ev.get(k)
        if side is None or ev is None:           # (forward ran without concurrency): plain execution
            fn()
            return
        if ev is not True:
            side.wait_event(ev)
            self._fork_ev[k] = True              # waited once: later ops of the section are ordered by the stream itself
        prev = self.stream
        with torch.cuda.stream(side):
            self.stream = side.cuda_stream
            self._exec_tag = self._tag = k       # (temporaries the op allocates belong to the section too)
            try:
                fn()
            finally:
                self._exec_tag = self._tag = 0
                self.stream = prev
        self._dirty.add(k)


# ---------------------------------------------------------------------------
# weight packing cache: (id(param), mode) -> [version, data_ptr, packed tensor, weakref, repacked-in-this-capture, generation]
#
# Validity = (tensor._version, _GEN).  `_version` catches in-place updates made through the tensor itself
# (optimizer.step() of torch.optim, load_state_dict, p.mul_()); it does NOT move for writes through `p.data`
# (e.g. the reference's radam.py: `p.data.copy_(p_data_fp32)`), so every optimizer step -- of ANY
# torch.optim.Optimizer subclass, via a global post-step hook -- also bumps the generation counter `_GEN`.
# Code that writes weights through `.data` outside an optimizer must call `invalidate_packed()` itself.
# A stale entry is re-packed IN PLACE: a captured CUDA graph keeps reading the buffer address it saw at capture.
_PACK = {}
_GEN = 0
FORCE_PACK = False      # set while a CUDA graph is being captured: pack kernels must be part of the graph


def invalidate_packed():
    """Mark every cached packed / tiled weight image stale (they are rebuilt, in place, on next use)."""
    global _GEN
    _GEN += 1


def _optimizer_step_hook(optimizer, args, kwargs):
    invalidate_packed()


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_hook
    _reg_hook(_optimizer_step_hook)
except ImportError:            # very old torch: callers must invalidate by hand
    pass


def weight_generation(w):
    return (w._version, _GEN)


def _cache_lookup(tp, key, w, numel, repack):
    """Shared cache protocol -> packed tensor.  `repack(dst_ptr)` enqueues the pack kernel(s) into dst."""
    ent = _PACK.get(key)
    if ent is not None and ent[3]() is w and ent[1] == w.data_ptr() and ent[2].numel() == numel:
        fresh = ent[0] == w._version and ent[5] == _GEN
        if (FORCE_PACK and not ent[4]) or not (fresh or FORCE_PACK):
            repack(ent[2].data_ptr())           # in place: the address is what a captured graph reads
            ent[0], ent[5] = w._version, _GEN
            if FORCE_PACK:
                ent[4] = True
                tp.repacked.append(ent)
        return ent[2]
    out = torch.empty(numel, dtype=torch.float32, device=w.device)
    repack(out.data_ptr())
    _PACK[key] = [w._version, w.data_ptr(), out, weakref.ref(w, lambda _r, k=key: _PACK.pop(k, None)), False, _GEN]
    return out


def reset_capture_flags():
    """Clear the repacked-during-capture marks (GraphedStep calls this in a finally: a capture that recorded a
    forward without its backward must not leave entries flagged)."""
    for ent in _PACK.values():
        ent[4] = False


def packed(tp, w, mode, A=None, Bc=None):
    """Packed GEMM layout of a conv / conv-transpose weight (see saunet_pack_weights)."""
    if not w.is_contiguous():
        raise RuntimeError("saunet_b200: conv weights must be contiguous")
    a, b, kh, kw = w.shape

    def repack(dst):
        _C.call("saunet_pack_weights", w.data_ptr(), dst, a, b, kh, kw, mode, tp.stream)
    return _cache_lookup(tp, (id(w), mode), w, w.numel(), repack).data_ptr()


# ---------------------------------------------------------------------------
# arithmetic class of the convolutions:
#   "fp32"   exact fp32 FFMA implicit GEMM (conv_simt.cu)
#   "3xtf32" tcgen05 tensor cores, hi/lo tf32 split of both operands, fp32 accumulate in TMEM (fp32-class accuracy)
#   "tf32"   tcgen05, single pass (what cuDNN does by default for PyTorch convs); ~1e-3 relative error
#   "bf16"   tcgen05 kind::f16 with bf16 operands and fp32 accumulation for every forward / data-gradient convolution
#            (activations, BatchNorm statistics, loss, master weights and gradients stay fp32 in memory; operands are rounded
#            to bf16 on their way into shared memory, weights are pre-tiled as bf16 images); weight gradients run
#            single-pass TF32 (their MN-major operand layouts exist for 32-bit elements only) -- BASELINE configs[2]
_PRECISION = os.environ.get("SAUNET_PRECISION", "3xtf32")
_WIDE_TILES = os.environ.get("SAUNET_WIDE_TILES", "1") == "1"


def set_precision(name):
    global _PRECISION
    if name not in ("fp32", "3xtf32", "tf32", "bf16"):
        raise ValueError("precision must be fp32 | 3xtf32 | tf32 | bf16")
    _PRECISION = name


def get_precision():
    return _PRECISION


def packed_tc(tp, w, mode, taps, Cin, N, phase=0, M=None, cm=False, wide=False):
    """Tensor-core tiling of a packed [K][N] weight (see saunet_pack_weights_tc) -> (ptr, BN, passes) or None.
    M (GEMM rows) lets small problems take a narrower N tile so that at least ~one CTA per SM exists."""
    if _PRECISION == "fp32":
        return None
    passes = {"3xtf32": 3, "tf32": 1, "bf16": 16}[_PRECISION]
    lib = _C.load()
    bn = lib.saunet_tc_tile_n(N)
    if wide:
        bn = 128
    elif M is not None:
        mt = (M + 127) // 128
        # persistent kernels: a 128-wide tile runs at ~2x the efficiency of the narrow ones (MMA operand traffic vs
        # math balance at N = 128), so keep it as long as most SMs get a tile; only really small problems trade tile
        # width for CTA count
        floor = 96 if bn >= 128 and _WIDE_TILES else 148
        while bn > 32 and mt * ((N + bn - 1) // bn) < floor:
            bn //= 2
            floor = 148
    K = taps * Cin

    def repack(dst):
        kn = packed(tp, w, mode) + 4 * phase * K * N
        _C.call("saunet_pack_weights_tc_cm" if cm else "saunet_pack_weights_tc", kn, taps, Cin, N, bn, passes, dst, tp.stream)
    numel = lib.saunet_tc_packed_floats_cm(taps, Cin, N, bn, passes) if cm else lib.saunet_tc_packed_floats(K, N, bn, passes)
    out = _cache_lookup(tp, (id(w), mode, phase, bn, passes, cm), w, numel, repack)
    return out.data_ptr(), bn, passes, 1 if cm else 0


def _wants_cm(x, KH, KW, stride, pad):
    """Mirror of conv_halo_tma_eligible (csrc/conv_halo_tma.cu) for channel counts that are not a multiple of 32: those
    3x3 layers (res3: 16 channels, dec1: 48) reach the TMA-fed kernel through the chunk-major padded weight image."""
    return KH == 3 and KW == 3 and stride == 1 and pad == 1 and x.H % 16 == 0 and x.W % 8 == 0 and x.C % 32 != 0


def _tc_ok(x, Cout, K):
    return _PRECISION != "fp32" and x.C % 4 == 0 and x.ld % 4 == 0 and x.ptr % 16 == 0 and Cout >= 8 and K >= 32


def _p(t):
    return 0 if t is None else t.data_ptr()


# ---------------------------------------------------------------------------
# thin op wrappers
def conv(tp, x, wptr, Cout, KH, KW, y, Hg, Wg, sy=1, sx=1, offy=0, offx=0, osy=1, osx=1, oy0=0, ox0=0,
         pro=0, pro_relu=0, bias=0, row_scale=0, row_add=0.0, act=ACT_NONE, acc=0, stat=None, wtc=None, epi=None):
    d = ConvDesc()
    d.x, d.x_ld, d.B, d.Hin, d.Win, d.Cin = x.ptr, x.ld, x.B, x.H, x.W, x.C
    d.w, d.Cout, d.KH, d.KW = wptr, Cout, KH, KW
    d.Hg, d.Wg, d.sy, d.sx, d.offy, d.offx = Hg, Wg, sy, sx, offy, offx
    d.y, d.y_ld, d.Hout, d.Wout, d.osy, d.osx, d.oy0, d.ox0 = y.ptr, y.ld, y.H, y.W, osy, osx, oy0, ox0
    if pro:
        d.in_scale, d.in_shift, d.in_relu = pro, pro + 4 * x.C, pro_relu
    else:
        d.in_scale, d.in_shift, d.in_relu = None, None, 0
    d.bias = bias or None
    d.row_scale, d.row_scale_add = row_scale or None, row_add
    d.act, d.accumulate = act, acc
    if stat is not None:
        d.stat_sum, d.stat_sumsq = stat
    else:
        d.stat_sum, d.stat_sumsq = None, None
    if wtc is not None:
        d.w_tc, d.tc_bn, d.tc_passes = wtc[:3]
        d.tc_cm = wtc[3] if len(wtc) > 3 else 0
    else:
        d.w_tc, d.tc_bn, d.tc_passes, d.tc_cm = None, 0, 0, 0
    if epi is not None:        # (xbuf, bn_state_ptr, relu): fused BatchNorm-backward epilogue, see bn_dgrad_fused
        xb, st, relu = epi
        d.epi_x, d.epi_x_ld, d.epi_relu = xb.ptr, xb.ld, relu
        d.epi_scale, d.epi_shift, d.epi_mean = st, st + 4 * xb.C, st + 8 * xb.C
    else:
        d.epi_x, d.epi_x_ld, d.epi_scale, d.epi_shift, d.epi_mean, d.epi_relu = None, 0, None, None, None, 0
    M = x.B * Hg * Wg
    _C.call("saunet_conv2d_fwd", ctypes.byref(d), tp.stream, flops=2.0 * M * KH * KW * x.C * Cout,
            nbytes=4.0 * (x.npix * x.C + M * Cout + KH * KW * x.C * Cout),
            tag="M%d K%dx%dx%d N%d s%d%s%s%s" % (M, KH, KW, x.C, Cout, sy, " pro" if pro else "", " tc" if wtc else "",
                                                   " bnbwd" if epi is not None else ""))


def wgrad(tp, p, q, dwptr, KH, KW, Hg, Wg, sy=1, sx=1, offy=0, offx=0, pro=0, pro_relu=0):
    d = WgradDesc()
    d.p, d.p_ld, d.Ca = p.ptr, p.ld, p.C
    d.q, d.q_ld, d.Cb, d.B, d.Hq, d.Wq = q.ptr, q.ld, q.C, q.B, q.H, q.W
    d.KH, d.KW, d.Hg, d.Wg, d.sy, d.sx, d.offy, d.offx = KH, KW, Hg, Wg, sy, sx, offy, offx
    if pro:
        d.q_scale, d.q_shift, d.q_relu = pro, pro + 4 * q.C, pro_relu
    else:
        d.q_scale, d.q_shift, d.q_relu = None, None, 0
    d.dw = dwptr
    d.precision = {"fp32": 0, "3xtf32": 1, "tf32": 2, "bf16": 2}[_PRECISION]
    M = q.B * Hg * Wg
    _C.call("saunet_conv2d_wgrad", ctypes.byref(d), tp.stream, flops=2.0 * M * KH * KW * p.C * q.C,
            nbytes=4.0 * M * (p.C + q.C), tag="M%d taps%d Ca%d Cb%d s%d" % (M, KH * KW, p.C, q.C, sy))


def channel_stats(tp, x, sum_ptr, sumsq_ptr):
    _C.call("saunet_channel_stats", x.ptr, x.ld, x.C, x.npix, sum_ptr, sumsq_ptr, tp.stream)


def affine_act(tp, x, state, y, act=ACT_NONE, res=None):
    _C.call("saunet_affine_act", x.ptr, x.ld, state or None, (state + 4 * x.C) if state else None,
            res.ptr if res is not None else None, res.ld if res is not None else 0, y.ptr, y.ld, x.C, x.npix, act,
            tp.stream)


def act_bwd(tp, dy, y, dz, act):
    _C.call("saunet_act_bwd", dy.ptr, dy.ld, y.ptr, y.ld, dz.ptr, dz.ld, dy.C, dy.npix, act, tp.stream)


def copy_slice(tp, src, dst, acc=0):
    _C.call("saunet_copy_slice", src.ptr, src.ld, dst.ptr, dst.ld, src.C, src.npix, acc, tp.stream)


def bias_grad(tp, dy, bias_param):
    """d(bias)[c] += sum over pixels of dy[:, c] (fp64 column sums)."""
    s = tp.dzeros(2 * dy.C)
    channel_stats(tp, dy, s, s + 8 * dy.C)
    _C.call("saunet_add_d2f", s, tp.pgrad(bias_param), dy.C, tp.stream)


# ---------------------------------------------------------------------------
class BN:
    """One BatchNorm application: finalised state [scale, shift, mean, invstd] + what backward needs."""
    __slots__ = ("mod", "state", "C", "training", "count")


def bn_stats_slot(tp, C):
    """(sum_ptr, sumsq_ptr) for C channels, zeroed."""
    s = tp.dzeros(2 * C)
    return (s, s + 8 * C)


def bn_finalize(tp, mod, C, stat, count):
    """Finalise a BatchNorm layer (nn.BatchNorm2d semantics):
    training -> batch statistics from (sum, sumsq, count) + running-stat update;
    eval -> running statistics."""
    bn = BN()
    bn.mod, bn.C, bn.count = mod, C, count
    bn.training = bool(mod.training or not mod.track_running_stats)
    bn.state = tp.fempty(4 * C)
    upd = mod.training and mod.track_running_stats and mod.running_mean is not None
    mom = mod.momentum
    if upd and mom is None:
        mom = 1.0 / float(int(mod.num_batches_tracked) + 1)
    if bn.training:
        assert stat is not None
        _C.call("saunet_bn_finalize", stat[0], stat[1], float(count), _p(mod.weight), _p(mod.bias),
                mod.running_mean.data_ptr() if upd else None, mod.running_var.data_ptr() if upd else None,
                float(mom or 0.0), float(mod.eps), 1, C, bn.state, tp.stream)
        if upd and mod.num_batches_tracked is not None:
            tp.bn_tracked.append(mod.num_batches_tracked)
    else:
        _C.call("saunet_bn_finalize", None, None, 1.0, _p(mod.weight), _p(mod.bias), mod.running_mean.data_ptr(),
                mod.running_var.data_ptr(), 0.0, float(mod.eps), 0, C, bn.state, tp.stream)
    return bn


def bn_backward(tp, bn, dy, x, out, act, dx, dx_acc, dres=None, dres_acc=0):
    """Two-pass BN(+ReLU) backward.  ``out`` (post-activation tensor) supplies the ReLU mask when it was
    materialised, else the mask is recomputed from x.  dx may alias dy."""
    C = bn.C
    red = tp.dzeros(2 * C)
    nout = 1 if (out is not None and act == ACT_RELU) else 0
    _C.call("saunet_bn_bwd_reduce", dy.ptr, dy.ld, x.ptr, x.ld, out.ptr if out is not None else None,
            out.ld if out is not None else 0, bn.state, C, dy.npix, act, red, tp.stream,
            nbytes=4.0 * C * dy.npix * (2 + nout), tag="C%d npix%d" % (C, dy.npix))
    mod = bn.mod
    has_affine = mod.weight is not None
    _C.call("saunet_bn_bwd_apply", dy.ptr, dy.ld, x.ptr, x.ld, out.ptr if out is not None else None,
            out.ld if out is not None else 0, bn.state, _p(mod.weight), red, C, dy.npix, act, 1 if bn.training else 0,
            dx.ptr if dx is not None else None, dx.ld if dx is not None else 0, dx_acc,
            dres.ptr if dres is not None else None, dres.ld if dres is not None else 0, dres_acc,
            tp.pgrad(mod.weight) if has_affine else None, tp.pgrad(mod.bias) if has_affine else None, tp.stream,
            nbytes=4.0 * C * dy.npix * (2 + nout + (1 + dx_acc if dx is not None else 0) + (1 + dres_acc if dres is not None else 0)),
            tag="C%d npix%d" % (C, dy.npix))


# ---------------------------------------------------------------------------
# composite: a convolution layer (nn.Conv2d semantics, stride 1 or 2, "same"-style padding) with its backward
class ConvRec:
    __slots__ = ("x", "y", "w", "b", "k", "stride", "pad", "pro", "pro_relu")


def conv2d(tp, x, w, b, y=None, stride=1, pad=0, pro=None, pro_relu=0, act=ACT_NONE, stat=None, row_scale=0,
           row_add=0.0):
    """y = act(rowscale * (conv(prologue(x), w) + b)); returns (y, rec) -- rec feeds conv2d_bwd."""
    Cout, Cin, KH, KW = w.shape
    assert Cin == x.C, "conv2d: weight expects %d input channels, got %d" % (Cin, x.C)
    Ho = (x.H + 2 * pad - KH) // stride + 1
    Wo = (x.W + 2 * pad - KW) // stride + 1
    if y is None:
        y = tp.new(x.B, Ho, Wo, Cout)
    assert (y.H, y.W, y.C) == (Ho, Wo, Cout)
    K = KH * KW * Cin
    conv(tp, x, packed(tp, w, 0), Cout, KH, KW, y, Ho, Wo, sy=stride, sx=stride, offy=-pad, offx=-pad,
         pro=pro.state if pro is not None else 0, pro_relu=pro_relu, bias=_p(b), row_scale=row_scale, row_add=row_add,
         act=act, stat=stat, wtc=packed_tc(tp, w, 0, KH * KW, Cin, Cout, M=x.B * Ho * Wo,
                                           cm=_wants_cm(x, KH, KW, stride, pad)) if _tc_ok(x, Cout, K) else None)
    r = ConvRec()
    r.x, r.y, r.w, r.b, r.k, r.stride, r.pad, r.pro, r.pro_relu = x, y, w, b, (KH, KW), stride, pad, pro, pro_relu
    return y, r


def conv2d_bwd(tp, r, dy, dx=None, dx_acc=0, need_bias=True, async_wgrad=None):
    """Given dy = d loss / d (conv output incl. bias): weight grad, bias grad and (if dx given) data grad
    w.r.t. the PROLOGUE OUTPUT (i.e. the activated tensor the GEMM consumed)."""
    w, x = r.w, r.x
    Cout, Cin, KH, KW = w.shape
    if w.requires_grad:
        pk = tp.packed_grad(w)
        # the weight gradient only reads dy and x and writes its own packed image: small problems run it beside the data
        # gradient on a side stream (round-robin over two) instead of in front of it
        side = (dx is not None if async_wgrad is None else async_wgrad) and dy.npix <= ASYNC_WGRAD_MAX_PIX
        ctx = tp.on_side(1 + (tp._wg_rr & 1), hold=(dy.s, x.s)) if side else contextlib.nullcontext()
        if side:
            tp._wg_rr += 1
        with ctx:
            dwp = None if pk else torch.zeros(w.numel(), dtype=torch.float32, device=tp.device)
            wgrad(tp, dy, x, pk or dwp.data_ptr(), KH, KW, dy.H, dy.W, sy=r.stride, sx=r.stride, offy=-r.pad, offx=-r.pad,
                  pro=r.pro.state if r.pro is not None else 0, pro_relu=r.pro_relu)
            if not pk:
                _C.call("saunet_unpack_wgrad", dwp.data_ptr(), tp.pgrad(w), Cout, Cin, KH, KW, 1, tp.stream)
    if r.b is not None and r.b.requires_grad:
        if need_bias:
            bias_grad(tp, dy, r.b)
        else:
            tp.pgrad(r.b)             # analytically zero: the (zero-initialised) gradient still exists
    if dx is not None:
        if r.stride != 1:
            raise RuntimeError("saunet_b200: data gradient of a strided conv is not on the SAUNet path")
        Kd = KH * KW * Cout
        conv(tp, dy, packed(tp, w, 1), Cin, KH, KW, dx, x.H, x.W, offy=-(KH - 1 - r.pad), offx=-(KW - 1 - r.pad),
             acc=dx_acc, wtc=packed_tc(tp, w, 1, KH * KW, Cout, Cin, M=x.npix,
                                       cm=_wants_cm(dy, KH, KW, 1, r.pad)) if _tc_ok(dy, Cin, Kd) else None)


# ---- 1x1 data gradient fused with the backward of the BatchNorm(+ReLU) in front of the conv ------------------------
BN_FUSED = os.environ.get("SAUNET_BN_FUSED", "1") == "1"


def bn_dgrad_fused_ok(r, dy):
    """Can conv `r` (a 1x1 / stride-1 conv whose input is relu?(bn(x)) applied as a prologue) take the fused path?"""
    w = r.w
    return (BN_FUSED and r.pro is not None and r.k == (1, 1) and r.stride == 1 and r.pad == 0
            and _tc_ok(dy, w.shape[1], w.shape[0]) and r.x.C % 4 == 0)


def bn_dgrad_fused(tp, r, dy, gx, gx_acc):
    """d loss / d x for  y = conv1x1(relu?(bn(x)))  WITHOUT materialising d loss / d bn-output:

        gx (+)= scale * g,   g = (W^T dy) * relu-mask          -- epilogue of the data-gradient GEMM (conv_pw_t.cu)
        sums  = [sum g, sum g (x - mean)]                       -- same epilogue
        dgamma, dbeta, and the mean terms  ab[0][c] + ab[1][c] * x  still owed to gx  -- saunet_bn_fused_finish

    The owed terms accumulate per channel in ``gx.s.ab`` (one double[2][ld] per Store) and are subtracted by
    ``bn_fixup`` right before a channel range of gx is consumed.  In a DenseNet block every layer's norm1 reads ALL
    earlier channels: instead of a reduce pass (2 tensors) + an apply pass (4 tensors) over `cin` channels per layer on
    top of the GEMM's own output, the layer costs one read of x and one read-modify-write of gx inside the GEMM."""
    bn, x, w = r.pro, r.x, r.w
    Cout, Cin = w.shape[0], w.shape[1]
    sums = tp.dzeros(2 * Cin)
    conv(tp, dy, packed(tp, w, 1), Cin, 1, 1, gx, x.H, x.W, acc=gx_acc, stat=(sums, sums + 8 * Cin),
         wtc=packed_tc(tp, w, 1, 1, Cout, Cin, wide=True), epi=(x, bn.state, r.pro_relu))
    if bn.training and not gx.s.ab:
        gx.s.ab = tp.dzeros(2 * gx.s.ld)
    mod = bn.mod
    aff = mod.weight is not None
    _C.call("saunet_bn_fused_finish", sums, bn.state, float(bn.count), 1 if bn.training else 0, Cin,
            tp.pgrad(mod.weight) if aff else None, tp.pgrad(mod.bias) if aff else None,
            (gx.s.ab + 8 * gx.c0) if bn.training else None, gx.s.ld, tp.stream)


def bn_fixup(tp, g, x):
    """Subtract the BatchNorm mean terms still owed to the gradient slice ``g`` (of the activation slice ``x``)."""
    if g.s.ab:
        _C.call("saunet_bn_fixup", g.ptr, g.ld, x.ptr, x.ld, g.s.ab + 8 * g.c0, g.s.ld, g.C, g.npix, tp.stream,
                nbytes=12.0 * g.C * g.npix, tag="C%d npix%d" % (g.C, g.npix))


# ---- ConvTranspose2d k4 s2 p1 (attention_blocks.py:179-183, models.py:211) as 4 output phases ----------
def convT4(tp, x, w, b, y, stat=None):
    Cin, Cout, KH, KW = w.shape
    assert (KH, KW) == (4, 4) and Cin == x.C and y.C == Cout and y.H == 2 * x.H and y.W == 2 * x.W
    wp = packed(tp, w, 2)
    K = 4 * Cin
    for pa in range(2):
        for pb in range(2):
            ph = pa * 2 + pb
            conv(tp, x, wp + 4 * ph * (K * Cout), Cout, 2, 2, y, x.H, x.W, offy=pa - 1, offx=pb - 1, osy=2, osx=2,
                 oy0=pa, ox0=pb, bias=_p(b), stat=stat,
                 wtc=packed_tc(tp, w, 2, 4, Cin, Cout, phase=ph) if _tc_ok(x, Cout, K) else None)


def convT4_bwd(tp, x, w, b, dy, dx, dx_acc, need_bias=True):
    Cin, Cout, _, _ = w.shape
    if w.requires_grad:
        pk = tp.packed_grad(w)
        dwp = None if pk else torch.zeros(w.numel(), dtype=torch.float32, device=tp.device)
        wgrad(tp, x, dy, pk or dwp.data_ptr(), 4, 4, x.H, x.W, sy=2, sx=2, offy=-1, offx=-1)
        if not pk:
            _C.call("saunet_unpack_wgrad", dwp.data_ptr(), tp.pgrad(w), Cin, Cout, 4, 4, 1, tp.stream)
    if b is not None and b.requires_grad:
        if need_bias:
            bias_grad(tp, dy, b)
        else:
            tp.pgrad(b)
    if dx is not None:
        conv(tp, dy, packed(tp, w, 0), Cin, 4, 4, dx, x.H, x.W, sy=2, sx=2, offy=-1, offx=-1, acc=dx_acc,
             wtc=packed_tc(tp, w, 0, 16, Cout, Cin) if _tc_ok(dy, Cin, 16 * Cout) else None)


# ---- resampling / pooling ------------------------------------------------------------------------------
def bilinear(tp, x, y):
    _C.call("saunet_bilinear_fwd", x.ptr, x.ld, x.B, x.H, x.W, x.C, y.ptr, y.ld, y.H, y.W, tp.stream)

    def bwd():
        dy = tp.grad(y)
        if dy is None:
            return
        dx, acc = tp.gw(x)
        _C.call("saunet_bilinear_bwd", dy.ptr, dy.ld, x.B, x.H, x.W, x.C, dx.ptr, dx.ld, y.H, y.W, acc, tp.stream)
    tp.on_backward(bwd)
    return y


def to_nhwc(tp, t):
    """NCHW-shaped fp32 CUDA tensor -> Buf (zero-copy when already channels_last)."""
    B, C, H, W = t.shape
    if t.dtype != torch.float32:
        raise RuntimeError("saunet_b200: fp32 tensors expected, got %s" % t.dtype)
    v = t.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return tp.wrap(v.reshape(-1), B, H, W, C)
    tc = t.contiguous()
    out = tp.new(B, H, W, C)
    _C.call("saunet_nchw_to_nhwc", tc.data_ptr(), out.ptr, out.ld, B, C, H * W, tp.stream)
    tp._keep.append(tc)
    return out


def grad_to_nhwc(tp, g):
    """Incoming NCHW-shaped gradient -> contiguous NHWC flat tensor (zero-copy when channels_last)."""
    B, C, H, W = g.shape
    v = g.permute(0, 2, 3, 1)
    if v.is_contiguous() and g.dtype == torch.float32:
        return v.reshape(-1)
    gc = g.to(torch.float32).contiguous()
    out = torch.empty(B * H * W * C, dtype=torch.float32, device=g.device)
    _C.call("saunet_nchw_to_nhwc", gc.data_ptr(), out.data_ptr(), C, B, C, H * W, tp.stream)
    tp._keep.append(gc)
    return out


# ---------------------------------------------------------------------------
class _TapeFn(torch.autograd.Function):
    """The module boundary: forward runs ``body(tape, *input_bufs)`` -> list of output Bufs; backward seeds the
    output grads, replays the tape and hands back input and parameter gradients."""

    @staticmethod
    def forward(ctx, body, arena, record, n_in, *args):
        ins, params = args[:n_in], args[n_in:]
        dev = ins[0].device
        tp = Tape(dev, record)
        tp.arena = arena
        in_bufs = [to_nhwc(tp, t.detach()) for t in ins]
        outs = body(tp, *in_bufs)
        if tp.bn_tracked:
            torch._foreach_add_(tp.bn_tracked, 1)
        ctx.tp, ctx.in_bufs, ctx.outs, ctx.params = tp, in_bufs, outs, params
        ctx.in_need = list(ctx.needs_input_grad[4:4 + n_in])
        res = tuple(o.nchw() for o in outs)
        if not record:
            for ent in tp.repacked:
                ent[4] = False
            ctx.tp = None
            ctx.in_bufs = ctx.outs = None
        return res

    @staticmethod
    def backward(ctx, *gouts):
        tp = ctx.tp
        if tp is None:
            raise RuntimeError("saunet_b200: backward through a forward that did not record a tape")
        tp.stream = torch.cuda.current_stream(tp.device).cuda_stream
        if tp.arena is not None:
            tp.arena.ensure_attached()      # zero_grad(set_to_none=True) detached the p.grad views: zero + re-attach
        for o, g in zip(ctx.outs, gouts):
            if g is None:
                continue
            gt = grad_to_nhwc(tp, g)
            if o.full() and o.s.g is None:
                tp.seed(o, gt)
            else:
                gb, acc = tp.gw(o)
                copy_slice(tp, Buf(Store(gt, o.npix, o.C), 0, o.C, o.B, o.H, o.W), gb, acc)
        tp.backward()
        gin = []
        for b, need in zip(ctx.in_bufs, ctx.in_need):
            g = tp.grad(b) if need else None
            gin.append(g.nchw() if g is not None else None)
        gp = [tp.pgrads.get(p) if p.requires_grad else None for p in ctx.params]
        for ent in tp.repacked:
            ent[4] = False
        ctx.tp = ctx.in_bufs = ctx.outs = None
        return (None, None, None, None) + tuple(gin) + tuple(gp)


def run(module, body, inputs):
    """Run ``body`` under one tape with ``module``'s parameters as autograd leaves."""
    seen, params = set(), []
    for p in module.parameters():
        if id(p) not in seen:
            seen.add(id(p))
            params.append(p)
    for t in inputs:
        if not t.is_cuda:
            raise RuntimeError("saunet_b200: input tensors must be CUDA tensors; there is no CPU fallback")
    # (grad mode is off inside Function.forward, so decide here whether a backward tape is needed)
    record = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or any(t.requires_grad for t in inputs))
    arena = getattr(module, "_saunet_grad_arena", None)
    if arena is not None and record:
        # Gradients land in the flat arena (p.grad are views of it), so autograd needs no edge to the ~700 parameters:
        # a fresh zero-size leaf carries "requires grad" instead.  Besides sparing autograd 700 AccumulateGrad nodes
        # per step, this keeps CUDA-graph capture safe: a parameter's AccumulateGrad node is pinned to the stream of
        # the first graph that used it and survives as long as any earlier loss tensor is alive -- capturing a
        # backward through it then fails with cudaErrorStreamCaptureIsolation.
        anchor = torch.empty(0, dtype=torch.float32, device=inputs[0].device, requires_grad=True)
        extra = [anchor] + [p for p in params if p.requires_grad and arena.ptr(p) is None]
    else:
        extra = params
    return _TapeFn.apply(body, arena, record, len(inputs), *inputs, *extra)
