"""The convolution kernels (tcgen05 implicit GEMM in all its variants, and the exact-fp32 FFMA kernel) through the
C-ABI entry points saunet_conv2d_fwd / saunet_conv2d_wgrad, on every geometry class of the SAUNet path, against an
INDEPENDENT float64 reference (torch F.conv2d / F.unfold in double -- no kernel, packer or descriptor of this repo is
on the reference side).  3xTF32 must agree to 2e-5 normalised (it carries ~21 mantissa bits), fp32 FFMA to 1e-5,
single-pass TF32 to 3e-3, bf16 operands (BASELINE configs[2]: kind::f16, fp32 accumulate) to 1.5e-2."""
import pytest
import torch

from saunet_b200 import _C, engine
from saunet_b200.engine import ACT_NONE, ACT_RELU, ACT_SIGMOID, Tape, conv, packed, packed_tc

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
DEFAULT_PRECISION = engine.get_precision()


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _ref_conv_fwd(x_nhwc, w, bvec, state, rs, act, y0, acc, stride, pad):
    """float64 reference of saunet_conv2d_fwd's contract (include/saunet_b200.h) -> (out [n,Cout], pre-activation sums)."""
    xa = x_nhwc.double()
    if state is not None:
        C = xa.shape[-1]
        xa = torch.relu(xa * state[:C].double() + state[C:].double())
    pre = torch.nn.functional.conv2d(xa.permute(0, 3, 1, 2), w.double(), bvec.double() if bvec is not None else None,
                                     stride=stride, padding=pad).permute(0, 2, 3, 1)
    Cout = pre.shape[-1]
    pre = pre.reshape(-1, Cout)
    o = pre * ((rs.double() + 1.0)[:, None] if rs is not None else 1.0)
    o = torch.relu(o) if act == ACT_RELU else (torch.sigmoid(o) if act == ACT_SIGMOID else o)
    if acc:
        o = o + y0.double()
    return o, torch.cat([pre.sum(0), (pre * pre).sum(0)])


def _ref_wgrad(P, Q, k, stride, off, state):
    """float64 reference of saunet_conv2d_wgrad: dw[(ky,kx,cb)][ca] = sum_pix P[pix][ca] * Q[gather(pix,ky,kx)][cb]."""
    B, H, W, Ca = P.shape
    q = Q.double()
    if state is not None:
        Cb = q.shape[-1]
        q = torch.relu(q * state[:Cb].double() + state[Cb:].double())
    cols = torch.nn.functional.unfold(q.permute(0, 3, 1, 2), kernel_size=k, stride=stride, padding=-off)   # [B, Cb*k*k, L]
    Cb = q.shape[-1]
    assert cols.shape[-1] == H * W
    cols = cols.view(B, Cb, k * k, H * W)
    return torch.einsum("bctl,bla->tca", cols, P.double().reshape(B, H * W, Ca)).reshape(-1)


CASES = [
    # B, H, W, Cin, Cout, k, stride, pad, x_ld_extra, y_ld_extra, prologue, bias, stats, rowscale, act, acc
    (2, 16, 16, 64, 128, 1, 1, 0, 0, 0, True, False, True, False, ACT_NONE, 0),      # dense 1x1 + BN prologue + stats
    (2, 16, 16, 128, 32, 3, 1, 1, 0, 96, True, False, True, False, ACT_NONE, 0),     # dense 3x3 into a concat slice
    (1, 20, 12, 64, 64, 3, 1, 1, 0, 0, False, False, True, False, ACT_NONE, 0),      # BasicBlock conv, ragged M
    (2, 9, 7, 48, 40, 3, 1, 1, 16, 0, False, True, False, False, ACT_RELU, 0),       # odd sizes, Cout not pow2, x slice
    (2, 8, 8, 256, 512, 3, 1, 1, 0, 0, False, True, True, False, ACT_NONE, 0),       # deep K, two N tiles
    (1, 16, 16, 32, 32, 1, 1, 0, 4, 0, False, False, False, True, ACT_NONE, 0),      # GSConv gate row scale
    (2, 12, 12, 16, 8, 1, 1, 0, 0, 0, False, True, False, False, ACT_SIGMOID, 0),    # K=16 < 32 padded k-block
    (2, 16, 16, 32, 64, 3, 1, 1, 0, 0, False, False, False, False, ACT_NONE, 1),     # accumulate (dgrad into grads)
    (1, 32, 32, 8, 16, 7, 2, 3, 0, 0, False, False, True, False, ACT_NONE, 0),       # 7x7 stride 2
    (3, 6, 5, 1024, 256, 1, 1, 0, 0, 0, True, False, False, False, ACT_NONE, 0),     # transition-like, K=1024
    # 3x3 / s1 / p1 with H % 16 == 0, W % 8 == 0, Cin % 32 == 0 -> the halo-patch kernel (conv_halo.cu)
    (2, 32, 24, 64, 64, 3, 1, 1, 0, 0, False, True, True, False, ACT_RELU, 0),
    (1, 16, 8, 32, 16, 3, 1, 1, 32, 16, True, False, True, False, ACT_NONE, 0),
    (3, 16, 40, 96, 128, 3, 1, 1, 0, 0, True, True, False, False, ACT_NONE, 1),
    (2, 48, 16, 160, 48, 3, 1, 1, 0, 0, False, False, True, False, ACT_SIGMOID, 0),
    # ... and channel counts that are not a multiple of 32: the TMA unit zero-fills the K padding (conv_halo_tma.cu)
    (2, 32, 16, 16, 16, 3, 1, 1, 0, 0, True, False, True, False, ACT_NONE, 0),       # res3: 16 -> 16 with BN prologue
    (1, 16, 24, 48, 64, 3, 1, 1, 16, 0, False, False, False, False, ACT_NONE, 1),    # dec1 dgrad: 48 -> 64, accumulate
    (2, 16, 8, 80, 32, 3, 1, 1, 0, 32, True, True, True, False, ACT_RELU, 0),        # 2.5 chunks
    (5, 64, 64, 32, 32, 3, 1, 1, 0, 0, False, False, True, False, ACT_NONE, 0),      # 160 tiles: persistent loop wraps, one chunk
    (2, 64, 48, 128, 32, 3, 1, 1, 0, 0, True, False, True, False, ACT_NONE, 0),      # dense conv2, several tiles per CTA
    (1, 32, 32, 64, 128, 3, 1, 1, 0, 0, False, True, False, False, ACT_NONE, 0),     # wide tile (two patch buffers)
    # large pointwise layers (>= 74 tiles of 256 pixels x 128 channels): also the shapes of the channels-on-lanes kernel
    # (conv_pw_t.cu)
    (2, 128, 128, 96, 128, 1, 1, 0, 32, 0, True, False, True, False, ACT_NONE, 0),   # dense conv1: prologue + statistics
    (1, 150, 150, 64, 160, 1, 1, 0, 0, 0, False, True, True, False, ACT_RELU, 0),    # ragged M, two channel tiles (one partial), bias
    (1, 160, 160, 128, 256, 1, 1, 0, 0, 64, False, False, False, False, ACT_NONE, 1),  # dgrad-like: accumulate into a slice
    (1, 192, 192, 40, 128, 1, 1, 0, 0, 0, True, True, False, True, ACT_SIGMOID, 0),  # K padded to 64, row gate, sigmoid
]


PREC_OF_PASSES = {3: "3xtf32", 1: "tf32", 16: "bf16"}


@pytest.mark.parametrize("passes,tol", [(3, 2e-5), (1, 3e-3), (16, 1.5e-2)])
@pytest.mark.parametrize("case", CASES)
def test_conv_matches_fp64(case, passes, tol):
    B, H, W, Cin, Cout, k, stride, pad, xe, ye, pro, bias, stats, rowscale, act, acc = case
    g = torch.Generator(device="cpu").manual_seed(hash(case) & 0xFFFF)
    tp = Tape(DEV, False)
    x = tp.new(B, H, W, Cin, ld=Cin + xe)
    x.s.t.copy_(torch.randn(x.s.t.numel(), generator=g).to(DEV))
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(DEV)
    w._version  # plain tensor: has a version counter
    bvec = torch.randn(Cout, generator=g).to(DEV) if bias else None
    state = torch.cat([0.5 + torch.rand(Cin, generator=g), 0.3 * torch.randn(Cin, generator=g)]).to(DEV) if pro else None
    rs = torch.rand(B * Ho * Wo, generator=g).to(DEV) if rowscale else None
    y0 = torch.randn(B * Ho * Wo * (Cout + ye), generator=g).to(DEV)
    ref_out, ref_sums = _ref_conv_fwd(x.s.t.view(B, H, W, Cin + xe)[..., :Cin], w, bvec, state, rs, act,
                                      y0.view(-1, Cout + ye)[:, :Cout], acc, stride, pad)
    outs, sums = [], []
    for use_tc in (False, True):
        engine.set_precision("fp32" if not use_tc else PREC_OF_PASSES[passes])
        y = tp.new(B, Ho, Wo, Cout, ld=Cout + ye)
        y.s.t.copy_(y0)
        st = torch.zeros(2 * Cout, dtype=torch.float64, device=DEV)
        wtc = packed_tc(tp, w, 0, k * k, Cin, Cout, cm=engine._wants_cm(x, k, k, stride, pad)) if use_tc else None
        assert (wtc is not None) == use_tc
        conv(tp, x, packed(tp, w, 0), Cout, k, k, y, Ho, Wo, sy=stride, sx=stride, offy=-pad, offx=-pad,
             pro=state.data_ptr() if pro else 0, pro_relu=1 if pro else 0, bias=bvec.data_ptr() if bias else 0,
             row_scale=rs.data_ptr() if rowscale else 0, row_add=1.0, act=act, acc=acc,
             stat=(st.data_ptr(), st.data_ptr() + 8 * Cout) if stats else None, wtc=wtc)
        torch.cuda.synchronize()
        outs.append(y.s.t.clone())
        sums.append(st.clone())
    engine.set_precision(DEFAULT_PRECISION)
    e32 = _rel(outs[0].view(-1, Cout + ye)[:, :Cout], ref_out)
    etc = _rel(outs[1].view(-1, Cout + ye)[:, :Cout], ref_out)
    print("case", case[:8], "passes", passes, "err vs fp64: fp32", e32, "tc", etc)
    assert e32 < 1e-5
    assert etc < tol
    if ye:      # channels outside the written slice are untouched
        for o in outs:
            assert torch.equal(o.view(-1, Cout + ye)[:, Cout:], y0.view(-1, Cout + ye)[:, Cout:])
    if stats:
        assert _rel(sums[0], ref_sums) < 1e-5
        assert _rel(sums[1], ref_sums) < max(tol, 1e-5)


def test_conv_tc_convT_phase_and_dgrad_shapes():
    """ConvTranspose2d 4x4 s2 p1 forward (4 phases) and its stride-2 data gradient on the tensor-core path vs fp32."""
    from saunet_b200.engine import convT4, convT4_bwd
    g = torch.Generator().manual_seed(5)
    tp = Tape(DEV, False)
    Cin, Cout, B, H, W = 64, 32, 2, 6, 5
    w = torch.nn.Parameter((torch.randn(Cin, Cout, 4, 4, generator=g) / (4 * Cin) ** 0.5).to(DEV), requires_grad=False)
    b = torch.nn.Parameter(torch.randn(Cout, generator=g).to(DEV), requires_grad=False)
    x = tp.new(B, H, W, Cin)
    x.s.t.copy_(torch.randn(x.s.t.numel(), generator=g).to(DEV))
    dy = tp.new(B, 2 * H, 2 * W, Cout)
    dy.s.t.copy_(torch.randn(dy.s.t.numel(), generator=g).to(DEV))
    res = {}
    for prec in ("fp32", "3xtf32", "bf16"):
        engine.set_precision(prec)
        y = tp.new(B, 2 * H, 2 * W, Cout)
        convT4(tp, x, w, b, y)
        dx = tp.new(B, H, W, Cin)
        convT4_bwd(tp, x, w, b, dy, dx, 0)
        torch.cuda.synchronize()
        res[prec] = (y.s.t.clone(), dx.s.t.clone())
    engine.set_precision(DEFAULT_PRECISION)
    ref = torch.nn.functional.conv_transpose2d(x.nchw().cpu().double(), w.detach().cpu().double(), b.detach().cpu().double(),
                                               stride=2, padding=1)
    xd = x.nchw().cpu().double().requires_grad_(True)
    torch.nn.functional.conv_transpose2d(xd, w.detach().cpu().double(), b.detach().cpu().double(), stride=2,
                                         padding=1).backward(dy.nchw().cpu().double())
    for prec, tol in (("fp32", 1e-5), ("3xtf32", 2e-5), ("bf16", 1.5e-2)):
        assert _rel(res[prec][0].view(B, 2 * H, 2 * W, Cout).permute(0, 3, 1, 2).cpu(), ref) < tol, prec
        assert _rel(res[prec][1].view(B, H, W, Cin).permute(0, 3, 1, 2).cpu(), xd.grad) < tol, prec


WG_CASES = [
    # B, H, W, Ca(P), Cb(Q), k, stride(sy), off, q_prologue, p_ld_extra, q_ld_extra
    (2, 16, 16, 32, 128, 3, 1, -1, True, 0, 0),       # dense conv2: Q wide (M side), P = 32
    (2, 16, 16, 128, 64, 1, 1, 0, True, 0, 192),      # dense conv1: P wide -> swapped, Q is a slice of a concat buffer
    (1, 20, 12, 64, 64, 3, 1, -1, False, 0, 0),       # BasicBlock, M side padded to 128
    (2, 9, 7, 40, 48, 3, 1, -1, False, 8, 16),        # ragged everything
    (2, 8, 8, 512, 256, 3, 1, -1, False, 0, 0),       # several M / N tiles
    (3, 24, 24, 16, 8, 1, 1, 0, False, 0, 0),         # tiny channel counts
    (2, 6, 5, 64, 32, 4, 2, -1, False, 0, 0),         # conv-transpose wgrad: P = x (low-res), Q = dY gathered at stride 2
    # pointwise convs -> conv_wgrad_pw.cu (dY on the 128 TMEM lanes, up to 512 Q channels per CTA)
    (2, 32, 32, 128, 224, 1, 1, 0, True, 0, 32),      # dense conv1, Q = first 224 channels of a 256-wide concat buffer
    (1, 16, 16, 128, 992, 1, 1, 0, True, 0, 32),      # deepest dense layer: two Q tiles of 496 channels
    (2, 24, 24, 256, 512, 1, 1, 0, True, 0, 0),       # transition: two dY tiles
    (1, 40, 40, 32, 64, 1, 1, 0, False, 0, 0),        # shape stream d1
    (3, 9, 7, 48, 40, 1, 1, 0, False, 16, 8),         # ragged pixel count (189 pixels), partial planes on both sides
    (2, 64, 64, 8, 32, 1, 1, 0, False, 0, 0),         # tiny dY
    (2, 16, 16, 64, 4, 7, 2, -3, False, 0, 0),        # the 7x7 / stride 2 stem on its 4-channel padded input
    # 3x3 / s1 / p1 with H % 8 == 0, W % 8 == 0, Ca <= 64 -> the all-taps halo kernel (conv_wgrad_halo.cu)
    (2, 16, 24, 32, 128, 3, 1, -1, True, 96, 0),      # dense conv2: dY is a slice of the concat-buffer gradient
    (3, 8, 8, 32, 128, 3, 1, -1, True, 0, 0),         # one tile per image (all-border halo)
    (1, 32, 16, 64, 64, 3, 1, -1, True, 0, 0),        # BasicBlock: 64-wide dY -> two tap groups
    (2, 16, 16, 32, 32, 3, 1, -1, False, 0, 0),       # res2: one Q plane
    (2, 24, 8, 16, 16, 3, 1, -1, True, 0, 0),         # res3: half planes on both sides
    (1, 16, 16, 48, 64, 3, 1, -1, False, 0, 0),       # dec1 3x3: Ca = 48
    (2, 16, 16, 32, 160, 3, 1, -1, False, 0, 32),     # two M tiles, the second partial
    (4, 128, 128, 32, 64, 3, 1, -1, False, 0, 0),     # dec0-like; 7 tiles per CTA (stage ring wraps around)
    (2, 128, 64, 64, 64, 3, 1, -1, True, 0, 0),       # res1-like; 4 tiles per CTA in each tap group
]


@pytest.mark.parametrize("case", WG_CASES)
def test_wgrad_matches_fp64(case):
    from saunet_b200.engine import wgrad
    B, H, W, Ca, Cb, k, stride, off, pro, pe, qe = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    tp = Tape(DEV, False)
    # output grid = P's pixel grid (H x W); Q lives on the gathered grid
    Hq, Wq = (H * stride, W * stride) if stride == 2 else (H, W)
    P = tp.new(B, H, W, Ca, ld=Ca + pe)
    P.s.t.copy_(torch.randn(P.s.t.numel(), generator=g).to(DEV))
    Q = tp.new(B, Hq, Wq, Cb, ld=Cb + qe)
    Q.s.t.copy_(torch.randn(Q.s.t.numel(), generator=g).to(DEV))
    state = torch.cat([0.5 + torch.rand(Cb, generator=g), 0.3 * torch.randn(Cb, generator=g)]).to(DEV) if pro else None
    res = []
    for prec in ("fp32", "3xtf32", "bf16"):          # (bf16 mode: weight gradients run single-pass TF32)
        engine.set_precision(prec)
        dw = torch.zeros(k * k * Cb * Ca, dtype=torch.float32, device=DEV)
        wgrad(tp, P, Q, dw.data_ptr(), k, k, H, W, sy=stride, sx=stride, offy=off, offx=off,
              pro=state.data_ptr() if pro else 0, pro_relu=1 if pro else 0)
        torch.cuda.synchronize()
        res.append(dw)
    engine.set_precision(DEFAULT_PRECISION)
    ref = _ref_wgrad(P.s.t.view(B, H, W, Ca + pe)[..., :Ca], Q.s.t.view(B, Hq, Wq, Cb + qe)[..., :Cb], k, stride, off, state)
    e32, etc, e1 = _rel(res[0], ref), _rel(res[1], ref), _rel(res[2], ref)
    print("wgrad case", case, "err vs fp64: fp32", e32, "3xtf32", etc, "tf32", e1)
    assert e32 < 1e-5
    assert etc < 2e-5
    assert e1 < 3e-3


SKINNY = [
    # B, H, W, Cin, Cout, x_ld_extra, y_ld_extra, prologue, bias, stats, rowscale, act, acc
    (2, 128, 128, 33, 33, 3, 0, True, True, False, False, ACT_RELU, 0),     # GSConv gate conv on a 36-wide cat buffer
    (2, 128, 128, 33, 1, 3, 0, False, True, True, False, ACT_NONE, 0),      # gate head + BN statistics
    (1, 256, 128, 8, 1, 0, 0, False, False, False, False, ACT_SIGMOID, 0),  # fuse
    (1, 128, 128, 2, 1, 0, 0, False, False, False, False, ACT_SIGMOID, 0),  # cw
    (1, 128, 128, 1, 32, 0, 32, False, True, True, False, ACT_NONE, 0),     # expand into a concat slice
    (2, 128, 128, 32, 4, 0, 0, False, True, False, False, ACT_NONE, 0),     # final
    (2, 128, 128, 4, 32, 0, 0, False, False, False, False, ACT_NONE, 1),    # final's data gradient, accumulating
    (2, 128, 128, 16, 16, 1, 0, False, False, False, True, ACT_NONE, 0),    # gated conv with the row gate
]


@pytest.mark.parametrize("case", SKINNY)
def test_skinny_conv_matches_reference(case):
    """conv_skinny.cu (one thread per pixel) against a float64 matmul reference, forward and weight gradient."""
    from saunet_b200.engine import wgrad
    B, H, W, Cin, Cout, xe, ye, pro, bias, stats, rowscale, act, acc = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    tp = Tape(DEV, False)
    n = B * H * W
    x = tp.new(B, H, W, Cin, ld=Cin + xe)
    x.s.t.copy_(torch.randn(x.s.t.numel(), generator=g).to(DEV))
    w = (torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5).to(DEV)
    bvec = torch.randn(Cout, generator=g).to(DEV) if bias else None
    state = torch.cat([0.5 + torch.rand(Cin, generator=g), 0.3 * torch.randn(Cin, generator=g)]).to(DEV) if pro else None
    rs = torch.rand(n, generator=g).to(DEV) if rowscale else None
    y = tp.new(B, H, W, Cout, ld=Cout + ye)
    y0 = torch.randn(y.s.t.numel(), generator=g).to(DEV)
    y.s.t.copy_(y0)
    st = torch.zeros(2 * Cout, dtype=torch.float64, device=DEV)
    conv(tp, x, packed(tp, w, 0), Cout, 1, 1, y, H, W, pro=state.data_ptr() if pro else 0, pro_relu=1 if pro else 0,
         bias=bvec.data_ptr() if bias else 0, row_scale=rs.data_ptr() if rowscale else 0, row_add=1.0, act=act, acc=acc,
         stat=(st.data_ptr(), st.data_ptr() + 8 * Cout) if stats else None)
    xa = x.s.t.view(n, Cin + xe)[:, :Cin].double()
    if pro:
        xa = torch.relu(xa * state[:Cin].double() + state[Cin:].double())
    pre = xa @ w.view(Cout, Cin).double().t() + (bvec.double() if bias else 0)
    o = pre * ((rs.double() + 1.0)[:, None] if rowscale else 1.0)
    o = torch.relu(o) if act == ACT_RELU else (torch.sigmoid(o) if act == ACT_SIGMOID else o)
    if acc:
        o = o + y0.view(n, Cout + ye)[:, :Cout].double()
    got = y.s.t.view(n, Cout + ye)
    assert _rel(got[:, :Cout], o) < 2e-6
    if ye:
        assert torch.equal(got[:, Cout:], y0.view(n, Cout + ye)[:, Cout:])
    if stats:
        assert _rel(st[:Cout], pre.sum(0)) < 1e-5 and _rel(st[Cout:], (pre * pre).sum(0)) < 1e-5
    # weight gradient with P = a random dY, Q = x (with the prologue)
    P = tp.new(B, H, W, Cout)
    P.s.t.copy_(torch.randn(P.s.t.numel(), generator=g).to(DEV))
    dw = torch.zeros(Cin * Cout, device=DEV)
    wgrad(tp, P, x, dw.data_ptr(), 1, 1, H, W, pro=state.data_ptr() if pro else 0, pro_relu=1 if pro else 0)
    ref = xa.t() @ P.s.t.view(n, Cout).double()
    assert _rel(dw.view(Cin, Cout), ref) < 1e-5
