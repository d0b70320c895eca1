"""Bandwidth-bound pointwise kernels through the C-ABI against torch fp64 references: align-corners bilinear resize
forward / backward (models/models.py:337-389; incl. the warp-parallel path for 1-channel maps upsampled x8 / x16) and
the GSConv gate backward (GSConv.py:55: out = conv(x * (alpha + 1)))."""
import pytest
import torch

from saunet_b200 import _C
from saunet_b200.engine import Tape

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("B,C,Hin,Hout,acc", [(2, 1, 16, 256, 0), (2, 1, 32, 256, 1), (1, 64, 64, 256, 0), (2, 8, 24, 48, 1),
                                               (1, 3, 20, 20, 0)])
def test_bilinear_fwd_bwd(B, C, Hin, Hout, acc):
    g = torch.Generator().manual_seed(B * 1000 + C * 10 + Hin)
    tp = Tape(DEV, False)
    x = torch.randn(B, Hin, Hin, C, generator=g).to(DEV)
    y = torch.empty(B, Hout, Hout, C, device=DEV)
    _C.call("saunet_bilinear_fwd", x.data_ptr(), C, B, Hin, Hin, C, y.data_ptr(), C, Hout, Hout, tp.stream)
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    yr = torch.nn.functional.interpolate(xr, size=(Hout, Hout), mode="bilinear", align_corners=True)
    assert _rel(y.permute(0, 3, 1, 2), yr.detach()) < 1e-5          # fp32 interpolation weights vs the fp64 reference
    dy = torch.randn(B, Hout, Hout, C, generator=g).to(DEV)
    dx0 = torch.randn(B, Hin, Hin, C, generator=g).to(DEV)
    dx = dx0.clone()
    _C.call("saunet_bilinear_bwd", dy.data_ptr(), C, B, Hin, Hin, C, dx.data_ptr(), C, Hout, Hout, acc, tp.stream)
    yr.backward(dy.double().permute(0, 3, 1, 2))
    ref = xr.grad.permute(0, 2, 3, 1) + (dx0.double() if acc else 0)
    assert _rel(dx, ref) < 2e-5


@pytest.mark.parametrize("C,ld_extra,acc", [(32, 0, 0), (16, 4, 1), (8, 0, 0), (4, 0, 0), (12, 0, 0)])
def test_rowscale_bwd(C, ld_extra, acc):
    """du = dout * (alpha + 1);  dalpha (+)= sum_c dout * out / (alpha + 1)   (out = W (x * (alpha + 1)) is linear in the gate)."""
    g = torch.Generator().manual_seed(C)
    tp = Tape(DEV, False)
    n = 4099                                               # not a multiple of anything
    ld = C + ld_extra
    dout = torch.randn(n, ld, generator=g).to(DEV)
    out = torch.randn(n, ld, generator=g).to(DEV)
    alpha = torch.rand(n, generator=g).to(DEV)
    du = torch.zeros(n, ld, device=DEV)
    da0 = torch.randn(n, generator=g).to(DEV)
    da = da0.clone()
    _C.call("saunet_rowscale_bwd", dout.data_ptr(), ld, out.data_ptr(), ld, alpha.data_ptr(), C, n, du.data_ptr(), ld,
            da.data_ptr(), acc, tp.stream)
    a1 = (alpha.double() + 1)[:, None]
    assert _rel(du[:, :C], dout[:, :C].double() * a1) < 1e-6
    assert float(du[:, C:].abs().max()) == 0.0 if ld_extra else True
    ref = (dout[:, :C].double() * out[:, :C].double() / a1).sum(1) + (da0.double() if acc else 0)
    assert _rel(da, ref) < 1e-5
