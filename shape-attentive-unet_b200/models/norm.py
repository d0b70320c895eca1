"""models/norm.py:16-22 of the reference: ``Norm2d`` = cfg.MODEL.BNFUNC(C) (torch.nn.BatchNorm2d by default)."""
from config import cfg


def Norm2d(in_channels):
    layer = getattr(cfg.MODEL, "BNFUNC")
    return layer(in_channels)
