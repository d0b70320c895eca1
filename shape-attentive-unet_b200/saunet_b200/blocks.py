"""Forward + hand-written backward of every SAUNet building block, on the Tape engine.

Each ``*_body`` function runs the block's forward through libsaunet_b200.so and
registers ONE backward closure (fwd+bwd fused per block, no torch autograd
inside).  Reference semantics restated (paths relative to /root/reference):
  BasicBlock            models/resnet.py:30-59
  GatedSpatialConv2d    models/GSConv.py:38-57
  DualAttBlock/_MRF/SpatialAttentionBlock/SEModule   models/attention_blocks.py:28-57,145-238
  DecoderBlock, conv3x3_bn_relu                      models/models.py:118-123,203-237
  _DenseLayer/_DenseBlock/_Transition                torchvision densenet.py:31-133
"""
import torch
from torch.nn.modules.batchnorm import _BatchNorm

from . import _C
from .engine import (ACT_NONE, ACT_RELU, ACT_SIGMOID, Buf, Store, _round4, act_bwd, affine_act, bn_backward,
                     bn_dgrad_fused, bn_dgrad_fused_ok, bn_finalize, bn_fixup, bn_stats_slot, channel_stats, conv2d,
                     conv2d_bwd, convT4, convT4_bwd, copy_slice)


def _check_bn(mod):
    if not isinstance(mod, _BatchNorm):
        raise RuntimeError("saunet_b200: norm layer %s is not a torch BatchNorm; cfg.MODEL.BNFUNC overrides are "
                           "not supported by the CUDA path" % type(mod).__name__)
    return mod


def _stat(tp, mod, C):
    _check_bn(mod)
    return bn_stats_slot(tp, C) if (mod.training or not mod.track_running_stats) else None


# ---------------------------------------------------------------------------
def conv_op(tp, x, w, b, y=None, stride=1, pad=0, act=ACT_NONE, x_needs_grad=True):
    """A stand-alone nn.Conv2d (+ optional fused activation) with its backward."""
    y, r = conv2d(tp, x, w, b, y=y, stride=stride, pad=pad, act=act)

    def bwd():
        dy = tp.grad(y)
        if dy is None:
            return
        if act != ACT_NONE:
            dz = tp.new(y.B, y.H, y.W, y.C)
            act_bwd(tp, dy, y, dz, act)
        else:
            dz = dy
        if x_needs_grad:
            dx, acc = tp.gw(x)
            conv2d_bwd(tp, r, dz, dx, acc)
        else:
            conv2d_bwd(tp, r, dz, None)
    tp.on_backward(bwd)
    return y


def cat_copy(tp, src, dst):
    """dst (a channel slice of a concat buffer) = src; gradient flows back by slice copy."""
    copy_slice(tp, src, dst)

    def bwd():
        g = tp.grad(dst)
        if g is None:
            return
        gs, acc = tp.gw(src)
        copy_slice(tp, g, gs, acc)
    tp.on_backward(bwd)
    return dst


def conv_bn_relu_body(tp, seq, x, out=None, pad=None):
    """nn.Sequential(Conv2d, BatchNorm2d, ReLU) (models/models.py:118-123; SAUNet.expand :299-301)."""
    cv, bnm = seq[0], _check_bn(seq[1])
    Co = cv.weight.shape[0]
    pad = cv.padding[0] if pad is None else pad
    st = _stat(tp, bnm, Co)
    t, r = conv2d(tp, x, cv.weight, cv.bias, pad=pad, stat=st)
    bn = bn_finalize(tp, bnm, Co, st, t.npix)
    if out is None:
        out = tp.new(t.B, t.H, t.W, Co)
    affine_act(tp, t, bn.state, out, ACT_RELU)

    def bwd():
        dout = tp.grad(out)
        if dout is None:
            return
        dt = tp.new(t.B, t.H, t.W, Co)
        bn_backward(tp, bn, dout, t, out, ACT_RELU, dt, 0)
        dx, acc = tp.gw(x)
        # a bias in front of a train-mode BatchNorm has an exactly zero gradient (the BN gradient sums to zero over the
        # batch): skip the extra column-sum pass over dt (the reference computes ~1e-7 of round-off there)
        conv2d_bwd(tp, r, dt, dx, acc, need_bias=not bn.training)
    tp.on_backward(bwd)
    return out


# ---------------------------------------------------------------------------
def basic_block_body(tp, m, x, out=None):
    if m.downsample is not None or m.stride != 1:
        raise NotImplementedError("saunet_b200 BasicBlock: only stride 1 without downsample is on the SAUNet path")
    C = m.conv1.weight.shape[0]
    if C != x.C:
        raise RuntimeError("BasicBlock: residual needs inplanes == planes (got %d -> %d)" % (x.C, C))
    n = x.npix
    st1 = _stat(tp, m.bn1, C)
    t1, r1 = conv2d(tp, x, m.conv1.weight, None, pad=1, stat=st1)
    bn1 = bn_finalize(tp, m.bn1, C, st1, n)
    st2 = _stat(tp, m.bn2, C)
    t2, r2 = conv2d(tp, t1, m.conv2.weight, None, pad=1, pro=bn1, pro_relu=1, stat=st2)
    bn2 = bn_finalize(tp, m.bn2, C, st2, n)
    if out is None:
        out = tp.new(x.B, x.H, x.W, C)
    affine_act(tp, t2, bn2.state, out, ACT_RELU, res=x)

    def bwd():
        dout = tp.grad(out)
        if dout is None:
            return
        dx, acc = tp.gw(x)
        dt2 = tp.new(x.B, x.H, x.W, C)
        bn_backward(tp, bn2, dout, t2, out, ACT_RELU, dt2, 0, dres=dx, dres_acc=acc)
        da1 = tp.new(x.B, x.H, x.W, C)
        conv2d_bwd(tp, r2, dt2, da1, 0)
        bn_backward(tp, bn1, da1, t1, None, ACT_RELU, da1, 0)
        conv2d_bwd(tp, r1, da1, dx, 1)
    tp.on_backward(bwd)
    return out


# ---------------------------------------------------------------------------
# GatedSpatialConv2d with C+1 gate channels padded to a multiple of 4.  The gate's (C+1)->(C+1) convolution is the
# block's only real arithmetic (2(C+1)^2 FLOP per pixel); with an odd channel count it can only run on the scalar FFMA
# kernels (257 -> 257 at 128x128, batch 32: 6.5 ms forward, 9.2 ms weight gradient).  Zero channels / zero weight rows and
# columns change nothing mathematically and put the three gate convolutions on the tensor-core kernels.
import os as _os
GS_PAD_MIN = int(_os.environ.get('SAUNET_GS_PAD_MIN', '41'))          # pad when C+1 >= this (smaller gates: the per-pixel skinny kernels are as fast)


class _PadBN:
    """Stands in for a BatchNorm module over the padded channel count (what bn_finalize / bn_backward read)."""

    def __init__(self, real, Cp):
        dev = real.running_mean.device
        self.real, self.Cp, self.C = real, Cp, real.num_features
        self.weight = torch.ones(Cp, device=dev)
        self.bias = torch.zeros(Cp, device=dev)
        self.running_mean = torch.zeros(Cp, device=dev)
        self.running_var = torch.ones(Cp, device=dev)

    training = property(lambda self: self.real.training)
    momentum = property(lambda self: self.real.momentum)
    eps = property(lambda self: self.real.eps)
    track_running_stats = property(lambda self: self.real.track_running_stats)
    num_batches_tracked = property(lambda self: self.real.num_batches_tracked)

    def pull(self):
        r, C = self.real, self.C
        with torch.no_grad():
            if r.weight is not None:
                self.weight[:C].copy_(r.weight)
                self.bias[:C].copy_(r.bias)
            self.running_mean[:C].copy_(r.running_mean)
            self.running_var[:C].copy_(r.running_var)

    def push_stats(self):
        with torch.no_grad():
            self.real.running_mean.copy_(self.running_mean[:self.C])
            self.real.running_var.copy_(self.running_var[:self.C])


def _gs_padded_params(m, Cp):
    """Persistent zero-padded copies of the gate's parameters, refreshed in place when the real ones moved."""
    from . import engine
    gc = m._gate_conv
    c1, c3 = gc[1], gc[3]
    C1 = c1.weight.shape[0]
    dev = c1.weight.device
    ent = getattr(m, "_saunet_pad", None)
    if ent is None or ent["dev"] != dev or ent["Cp"] != Cp:
        ent = {"dev": dev, "Cp": Cp, "gen": None, "bn0": _PadBN(gc[0], Cp),
               "w1": torch.zeros(Cp, Cp, 1, 1, device=dev).requires_grad_(True), "b1": torch.zeros(Cp, device=dev).requires_grad_(True),
               "w3": torch.zeros(1, Cp, 1, 1, device=dev).requires_grad_(True)}
        m._saunet_pad = ent
    gen = (engine.weight_generation(c1.weight), engine.weight_generation(c3.weight), c1.bias._version)
    if ent["gen"] != gen or engine.FORCE_PACK:
        with torch.no_grad():
            ent["w1"][:C1, :C1].copy_(c1.weight)
            ent["b1"][:C1].copy_(c1.bias)
            ent["w3"][:, :C1].copy_(c3.weight)
        ent["gen"] = gen
    ent["bn0"].pull()
    return ent


def _scatter_grad(tp, padded, real, rows, cols, cols_pad):
    """real.grad[rows, cols] += padded.grad[:rows, :cols] (row pitch cols_pad)."""
    pg = tp.pgrads.get(padded)
    if pg is not None and real is not None and real.requires_grad:
        _C.call("saunet_copy_slice", pg.data_ptr(), cols_pad, tp.pgrad(real), cols, cols, rows, 1, tp.stream)


def gsconv_body_padded(tp, m, x, g):
    C = x.C
    C1, Cp = C + 1, _round4(C + 1)
    n = x.npix
    B, H, W = x.B, x.H, x.W
    xg = tp.new(B, H, W, Cp)
    xg.s.t.zero_()                                  # the pad channels must be exactly zero
    copy_slice(tp, x, xg.slice(0, C))
    copy_slice(tp, g, xg.slice(C, 1))
    xin = xg.slice(0, C)
    gc = m._gate_conv
    c1, c3, bn4m = gc[1], gc[3], _check_bn(gc[4])
    _check_bn(gc[0])
    P = _gs_padded_params(m, Cp)
    bn0m, w1, b1, w3 = P["bn0"], P["w1"], P["b1"], P["w3"]
    st0 = bn_stats_slot(tp, Cp) if (bn0m.training or not bn0m.track_running_stats) else None
    if st0 is not None:
        channel_stats(tp, xg, st0[0], st0[1])
    bn0 = bn_finalize(tp, bn0m, Cp, st0, n)
    if st0 is not None and bn0m.track_running_stats:
        bn0m.push_stats()
    h1, r1 = conv2d(tp, xg, w1, b1, pro=bn0, pro_relu=0, act=ACT_RELU)
    st4 = _stat(tp, bn4m, 1)
    a_pre, r3 = conv2d(tp, h1, w3, c3.bias, stat=st4)
    bn4 = bn_finalize(tp, bn4m, 1, st4, n)
    alphas = tp.new(B, H, W, 1)
    affine_act(tp, a_pre, bn4.state, alphas, ACT_SIGMOID)
    out, rmain = conv2d(tp, xin, m.weight, None, row_scale=alphas.ptr, row_add=1.0)

    def bwd():
        dout, dal_ext = tp.grad(out), tp.grad(alphas)
        if dout is None and dal_ext is None:
            return
        dalpha = tp.new(B, H, W, 1)
        du = None
        if dout is not None:
            du = tp.new(B, H, W, C)
            _C.call("saunet_rowscale_bwd", dout.ptr, dout.ld, out.ptr, out.ld, alphas.ptr, C, n, du.ptr, du.ld,
                    dalpha.ptr, 0, tp.stream)
            if dal_ext is not None:
                copy_slice(tp, dal_ext, dalpha, 1)
        else:
            copy_slice(tp, dal_ext, dalpha, 0)
        act_bwd(tp, dalpha, alphas, dalpha, ACT_SIGMOID)
        bn_backward(tp, bn4, dalpha, a_pre, None, ACT_NONE, dalpha, 0)
        dh1 = tp.new(B, H, W, Cp)
        conv2d_bwd(tp, r3, dalpha, dh1, 0, need_bias=not bn4.training)
        act_bwd(tp, dh1, h1, dh1, ACT_RELU)
        da0 = tp.new(B, H, W, Cp)
        conv2d_bwd(tp, r1, dh1, da0, 0)
        gxg, acc = tp.gw(xg)
        bn_backward(tp, bn0, da0, xg, None, ACT_NONE, gxg, acc)
        if du is not None:
            conv2d_bwd(tp, rmain, du, gxg.slice(0, C), 1)
        # gradients of the padded copies -> the real parameters
        _scatter_grad(tp, w1, c1.weight, C1, C1, Cp)
        _scatter_grad(tp, b1, c1.bias, 1, C1, Cp)
        _scatter_grad(tp, w3, c3.weight, 1, C1, Cp)
        _scatter_grad(tp, bn0m.weight, bn0m.real.weight, 1, C1, Cp)
        _scatter_grad(tp, bn0m.bias, bn0m.real.bias, 1, C1, Cp)
        gx, a = tp.gw(x)
        copy_slice(tp, gxg.slice(0, C), gx, a)
        gg, a = tp.gw(g)
        copy_slice(tp, gxg.slice(C, 1), gg, a)
    tp.on_backward(bwd)
    return out, alphas


def gsconv_body(tp, m, x, g):
    """-> (out, alphas).  x: [.,C], g: [.,1].  Zero-copy when g is the channel right after x in one Store."""
    from . import engine
    if (x.C + 1) % 4 and x.C + 1 >= GS_PAD_MIN and engine.get_precision() != "fp32" and m.bias is None \
            and tuple(m.kernel_size) == (1, 1) and g.C == 1 and x.C % 4 == 0:
        return gsconv_body_padded(tp, m, x, g)
    if m.bias is not None or tuple(m.kernel_size) != (1, 1) or tuple(m.stride) != (1, 1) or tuple(m.padding) != (0, 0) \
            or tuple(m.dilation) != (1, 1) or m.groups != 1:
        raise NotImplementedError("saunet_b200 GatedSpatialConv2d: only the 1x1 / stride 1 / no-bias form used by "
                                  "SAUNet (models/models.py:295-297) is implemented")
    C = x.C
    if g.C != 1 or (g.B, g.H, g.W) != (x.B, x.H, x.W):
        raise RuntimeError("GatedSpatialConv2d: gating features must be [N,1,H,W] matching the input")
    n = x.npix
    B, H, W = x.B, x.H, x.W
    copied = not (x.s is g.s and g.c0 == x.c0 + C)
    if copied:
        xg = tp.new(B, H, W, C + 1, ld=_round4(C + 1))
        copy_slice(tp, x, xg.slice(0, C))
        copy_slice(tp, g, xg.slice(C, 1))
    else:
        xg = Buf(x.s, x.c0, C + 1, B, H, W)
    xin = xg.slice(0, C)
    gc = m._gate_conv
    bn0m, c1, c3, bn4m = _check_bn(gc[0]), gc[1], gc[3], _check_bn(gc[4])
    st0 = _stat(tp, bn0m, C + 1)
    if st0 is not None:
        channel_stats(tp, xg, st0[0], st0[1])
    bn0 = bn_finalize(tp, bn0m, C + 1, st0, n)
    h1, r1 = conv2d(tp, xg, c1.weight, c1.bias, pro=bn0, pro_relu=0, act=ACT_RELU)
    st4 = _stat(tp, bn4m, 1)
    a_pre, r3 = conv2d(tp, h1, c3.weight, c3.bias, stat=st4)
    bn4 = bn_finalize(tp, bn4m, 1, st4, n)
    alphas = tp.new(B, H, W, 1)
    affine_act(tp, a_pre, bn4.state, alphas, ACT_SIGMOID)
    out, rmain = conv2d(tp, xin, m.weight, None, row_scale=alphas.ptr, row_add=1.0)

    def bwd():
        dout, dal_ext = tp.grad(out), tp.grad(alphas)
        if dout is None and dal_ext is None:
            return
        dalpha = tp.new(B, H, W, 1)
        du = None
        if dout is not None:
            du = tp.new(B, H, W, C)
            _C.call("saunet_rowscale_bwd", dout.ptr, dout.ld, out.ptr, out.ld, alphas.ptr, C, n, du.ptr, du.ld,
                    dalpha.ptr, 0, tp.stream)
            if dal_ext is not None:
                copy_slice(tp, dal_ext, dalpha, 1)
        else:
            copy_slice(tp, dal_ext, dalpha, 0)
        act_bwd(tp, dalpha, alphas, dalpha, ACT_SIGMOID)
        bn_backward(tp, bn4, dalpha, a_pre, None, ACT_NONE, dalpha, 0)
        dh1 = tp.new(B, H, W, C + 1)
        conv2d_bwd(tp, r3, dalpha, dh1, 0, need_bias=not bn4.training)
        act_bwd(tp, dh1, h1, dh1, ACT_RELU)
        da0 = tp.new(B, H, W, C + 1)
        conv2d_bwd(tp, r1, dh1, da0, 0)
        gxg, acc = tp.gw(xg)
        bn_backward(tp, bn0, da0, xg, None, ACT_NONE, gxg, acc)
        if du is not None:
            conv2d_bwd(tp, rmain, du, gxg.slice(0, C), 1)
        if copied:
            gx, a = tp.gw(x)
            copy_slice(tp, gxg.slice(0, C), gx, a)
            gg, a = tp.gw(g)
            copy_slice(tp, gxg.slice(C, 1), gg, a)
    tp.on_backward(bwd)
    return out, alphas


# ---------------------------------------------------------------------------
def dual_att_body(tp, m, lo, skip, mcat=None):
    """-> (out, spatial).  ``mcat``: optional pre-allocated concat buffer whose first skip.C channels ARE skip."""
    sa, se = m.spatialAttn, m.channelAttn
    if sa.normalize_attn:
        raise NotImplementedError("SpatialAttentionBlock(normalize_attn=True) is dead code in the reference "
                                  "(attention_blocks.py:169-170 would raise NameError)")
    C0, C1 = lo.C, skip.C
    B, H, W = skip.B, skip.H, skip.W
    if (H, W) != (2 * lo.H, 2 * lo.W) or lo.B != B:
        raise RuntimeError("DualAttBlock: skip must be 2x the resolution of the low-res input")
    n = B * H * W
    copied = mcat is None
    if copied:
        mcat = tp.new(B, H, W, C1 + C0)
        copy_slice(tp, skip, mcat.slice(0, C1))
    ct, bnum = m.mrf.up[0], _check_bn(m.mrf.up[1])
    tu = tp.new(B, H, W, C0)
    stu = _stat(tp, bnum, C0)
    convT4(tp, lo, ct.weight, ct.bias, tu, stat=stu)
    bnu = bn_finalize(tp, bnum, C0, stu, n)
    up = mcat.slice(C1, C0)
    affine_act(tp, tu, bnu.state, up, ACT_RELU)
    c3, bncm = m.c3x3rb[0], _check_bn(m.c3x3rb[1])
    Co = c3.weight.shape[0]
    stc = _stat(tp, bncm, Co)
    tc, rc = conv2d(tp, mcat, c3.weight, c3.bias, pad=1, stat=stc)
    bnc = bn_finalize(tp, bncm, Co, stc, n)
    fused = tp.new(B, H, W, Co)
    affine_act(tp, tc, bnc.state, fused, ACT_RELU)
    # ---- attention tail (timed in isolation by bench.py --workload blocks through _C.SCOPE) ----
    # spatial attention: sigmoid(phi(relu(bn(down(fused)))))
    _C.SCOPE = "tail"
    Ca = sa.down.weight.shape[0]
    std = _stat(tp, _check_bn(sa.bn), Ca)
    td, rd = conv2d(tp, fused, sa.down.weight, sa.down.bias, stat=std)
    bnd = bn_finalize(tp, sa.bn, Ca, std, n)
    S, rphi = conv2d(tp, td, sa.phi.weight, sa.phi.bias, pro=bnd, pro_relu=1, act=ACT_SIGMOID)
    # channel attention (SE, reduction 16)
    gapv = tp.new(B, 1, 1, Co)
    _C.call("saunet_gap_fwd", fused.ptr, fused.ld, B, H * W, Co, gapv.ptr, tp.stream)
    z1, rf1 = conv2d(tp, gapv, se.fc1.weight, se.fc1.bias, act=ACT_RELU)
    cv, rf2 = conv2d(tp, z1, se.fc2.weight, se.fc2.bias, act=ACT_SIGMOID)
    out = tp.new(B, H, W, Co)
    _C.call("saunet_dualatt_combine_fwd", fused.ptr, fused.ld, S.ptr, cv.ptr, B, H * W, Co, out.ptr, out.ld, tp.stream)
    _C.SCOPE = ""

    def bwd():
        dout, dS_ext = tp.grad(out), tp.grad(S)
        if dout is None and dS_ext is None:
            return
        _C.SCOPE = "tail"
        dS = tp.new(B, H, W, 1)
        if dout is not None:
            dfused = tp.new(B, H, W, Co)
            dcv_t = torch.zeros(B * Co, dtype=torch.float32, device=tp.device)
            dcv = Buf(Store(dcv_t, B, Co), 0, Co, B, 1, 1)
            _C.call("saunet_dualatt_combine_bwd", dout.ptr, dout.ld, fused.ptr, fused.ld, S.ptr, cv.ptr, B, H * W, Co,
                    dfused.ptr, dfused.ld, 0, dS.ptr, dcv.ptr, tp.stream)
            if dS_ext is not None:
                copy_slice(tp, dS_ext, dS, 1)
            # SE path
            act_bwd(tp, dcv, cv, dcv, ACT_SIGMOID)
            dz1 = tp.new(B, 1, 1, z1.C)
            conv2d_bwd(tp, rf2, dcv, dz1, 0)
            act_bwd(tp, dz1, z1, dz1, ACT_RELU)
            dgap = tp.new(B, 1, 1, Co)
            conv2d_bwd(tp, rf1, dz1, dgap, 0)
            _C.call("saunet_gap_bwd", dgap.ptr, B, H * W, Co, dfused.ptr, dfused.ld, 1, tp.stream)
        else:
            dfused = Buf(Store(torch.zeros(n * Co, dtype=torch.float32, device=tp.device), n, Co), 0, Co, B, H, W)
            copy_slice(tp, dS_ext, dS, 0)
        # spatial path
        act_bwd(tp, dS, S, dS, ACT_SIGMOID)
        dad = tp.new(B, H, W, Ca)
        conv2d_bwd(tp, rphi, dS, dad, 0)
        bn_backward(tp, bnd, dad, td, None, ACT_RELU, dad, 0)
        conv2d_bwd(tp, rd, dad, dfused, 1)
        _C.SCOPE = ""
        # c3x3rb
        bn_backward(tp, bnc, dfused, tc, fused, ACT_RELU, dfused, 0)
        gm, acc = tp.gw(mcat)
        conv2d_bwd(tp, rc, dfused, gm, acc, need_bias=not bnc.training)
        # mrf.up
        dtu = tp.new(B, H, W, C0)
        bn_backward(tp, bnu, gm.slice(C1, C0), tu, up, ACT_RELU, dtu, 0)
        dlo, acc = tp.gw(lo)
        convT4_bwd(tp, lo, ct.weight, ct.bias, dtu, dlo, acc, need_bias=not bnu.training)
        if copied:
            gs, a = tp.gw(skip)
            copy_slice(tp, gm.slice(0, C1), gs, a)
    tp.on_backward(bwd)
    return out, S


# ---------------------------------------------------------------------------
def decoder_block_body(tp, m, x, out=None):
    seq = m.block
    if not isinstance(seq[1], torch.nn.ConvTranspose2d):
        raise NotImplementedError("saunet_b200 DecoderBlock: only is_deconv=True (the SAUNet default) is implemented")
    cbr, ct, bntm = seq[0], seq[1], _check_bn(seq[2])
    a1 = conv_bn_relu_body(tp, cbr, x)
    Co = ct.weight.shape[1]
    B, H, W = x.B, 2 * a1.H, 2 * a1.W
    t2 = tp.new(B, H, W, Co)
    st = _stat(tp, bntm, Co)
    convT4(tp, a1, ct.weight, ct.bias, t2, stat=st)
    bnt = bn_finalize(tp, bntm, Co, st, t2.npix)
    if out is None:
        out = tp.new(B, H, W, Co)
    affine_act(tp, t2, bnt.state, out, ACT_RELU)

    def bwd():
        dout = tp.grad(out)
        if dout is None:
            return
        dt2 = tp.new(B, H, W, Co)
        bn_backward(tp, bnt, dout, t2, out, ACT_RELU, dt2, 0)
        da1, acc = tp.gw(a1)
        convT4_bwd(tp, a1, ct.weight, ct.bias, dt2, da1, acc, need_bias=not bnt.training)
    tp.on_backward(bwd)
    return out


# ---------------------------------------------------------------------------
# DenseNet-121 pieces (torchvision densenet.py:31-133).  X is the block's preallocated concat buffer; the first
# c_in channels are filled and (training) their per-channel sums already sit in `sums`.  Per-channel batch
# statistics of a feature are computed ONCE by the producing conv's epilogue and reused by every later norm1.
def dense_block_body(tp, blk, X, c_in, sums):
    n = X.npix
    B, H, W = X.B, X.H, X.W
    cin = c_in

    def bwd_block_input():
        # (registered first = runs last) BatchNorm mean terms the block's layers still owe to its input channels
        gX = tp.grad(X)
        if gX is not None:
            bn_fixup(tp, gX.slice(0, c_in), X.slice(0, c_in))
    tp.on_backward(bwd_block_input)
    for layer in blk.children():
        n1, n2 = _check_bn(layer.norm1), _check_bn(layer.norm2)
        mid, gr = layer.conv1.weight.shape[0], layer.conv2.weight.shape[0]
        train1 = n1.training or not n1.track_running_stats
        if train1 and sums is None:
            raise RuntimeError("dense block: a BatchNorm is in training mode but the block is in eval mode")
        bn1 = bn_finalize(tp, n1, cin, sums if train1 else None, n)
        st2 = _stat(tp, n2, mid)
        xin = X.slice(0, cin)
        t1, r1 = conv2d(tp, xin, layer.conv1.weight, None, pro=bn1, pro_relu=1, stat=st2)
        bn2 = bn_finalize(tp, n2, mid, st2, n)
        ynew = X.slice(cin, gr)
        _, r2 = conv2d(tp, t1, layer.conv2.weight, None, y=ynew, pad=1, pro=bn2, pro_relu=1,
                       stat=(sums[0] + 8 * cin, sums[1] + 8 * cin) if sums is not None else None)

        def bwd(cin=cin, mid=mid, gr=gr, bn1=bn1, bn2=bn2, t1=t1, r1=r1, r2=r2, xin=xin):
            gX = tp.grad(X)
            if gX is None:
                return
            # this layer's 32 output channels were read by the norm1 of every later layer (and the transition):
            # settle the mean terms of those BatchNorm gradients before the slice is consumed
            bn_fixup(tp, gX.slice(cin, gr), X.slice(cin, gr))
            da2 = tp.new(B, H, W, mid)
            conv2d_bwd(tp, r2, gX.slice(cin, gr), da2, 0)
            bn_backward(tp, bn2, da2, t1, None, ACT_RELU, da2, 0)
            if bn_dgrad_fused_ok(r1, da2):
                conv2d_bwd(tp, r1, da2, None, async_wgrad=True)
                bn_dgrad_fused(tp, r1, da2, gX.slice(0, cin), 1)
            else:
                da1 = tp.new(B, H, W, cin)
                conv2d_bwd(tp, r1, da2, da1, 0)
                bn_backward(tp, bn1, da1, xin, None, ACT_RELU, gX.slice(0, cin), 1)
        tp.on_backward(bwd)
        cin += gr
    return cin


def transition_body(tp, tr, X, sums, out):
    nm = _check_bn(tr.norm)
    train = nm.training or not nm.track_running_stats
    bn = bn_finalize(tp, nm, X.C, sums if train else None, X.npix)
    tt, r = conv2d(tp, X, tr.conv.weight, None, pro=bn, pro_relu=1)
    _C.call("saunet_avgpool2_fwd", tt.ptr, tt.ld, tt.B, tt.H, tt.W, tt.C, out.ptr, out.ld, tp.stream)

    def bwd():
        dout = tp.grad(out)
        if dout is None:
            return
        dtt = tp.new(tt.B, tt.H, tt.W, tt.C)
        _C.call("saunet_avgpool2_bwd", dout.ptr, dout.ld, tt.B, tt.H, tt.W, tt.C, dtt.ptr, dtt.ld, 0, tp.stream)
        if bn_dgrad_fused_ok(r, dtt):
            conv2d_bwd(tp, r, dtt, None, async_wgrad=True)
            gX, acc = tp.gw(X)
            bn_dgrad_fused(tp, r, dtt, gX, acc)      # mean terms settled per slice by the dense block's backward
        else:
            da = tp.new(X.B, X.H, X.W, X.C)
            conv2d_bwd(tp, r, dtt, da, 0)
            gX, acc = tp.gw(X)
            bn_backward(tp, bn, da, X, None, ACT_RELU, gX, acc)
    tp.on_backward(bwd)
    return out


def maxpool2(tp, x):
    y = tp.new(x.B, x.H // 2, x.W // 2, x.C)
    _C.call("saunet_maxpool2_fwd", x.ptr, x.ld, x.B, x.H, x.W, x.C, y.ptr, y.ld, tp.stream)

    def bwd():
        dy = tp.grad(y)
        if dy is None:
            return
        dx, acc = tp.gw(x)
        _C.call("saunet_maxpool2_bwd", dy.ptr, dy.ld, x.ptr, x.ld, x.B, x.H, x.W, x.C, dx.ptr, dx.ld, acc, tp.stream)
    tp.on_backward(bwd)
    return y
