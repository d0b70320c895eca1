"""Global config mirror of the reference's config.py:53-71.

Only ``cfg.MODEL.BNFUNC`` is read on the hot path (models/norm.py:16-22 of the
reference): the factory ``Norm2d`` uses to build its normalisation layer.
"""
import torch


class AttrDict(dict):
    """Minimal attribute-access dict (the reference's AttrDict.py, host-side config only)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


cfg = AttrDict()
cfg.EPOCH = 0
cfg.MODEL = AttrDict()
cfg.MODEL.BN = "regularnorm"
cfg.MODEL.BNFUNC = torch.nn.BatchNorm2d
cfg.MODEL.BIGMEMORY = False
