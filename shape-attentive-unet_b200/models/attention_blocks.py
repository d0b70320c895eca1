"""Dual (spatial + channel) attention decoder block of SAUNet
(models/attention_blocks.py:28-57,145-238 of the reference).

The nn.Module classes own the parameters under the reference's state_dict keys; the arithmetic runs in
saunet_b200.blocks.dual_att_body (CUDA, fwd+bwd fused per block).
"""
import math

import torch.nn as nn

from saunet_b200 import engine
from saunet_b200.blocks import dual_att_body


def _init(module):
    """attention_blocks.py:40-48,155-163,187-197,222-230: conv / conv-transpose weights ~ N(0, sqrt(2/(k*k*C_out))),
    BatchNorm (1, 0)."""
    for m in module.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
            m.weight.data.normal_(0, math.sqrt(2.0 / n))
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


class SEModule(nn.Module):
    """Channel attention: x * sigmoid(fc2(relu(fc1(avgpool(x)))))."""

    def __init__(self, channels, reduction):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, channels // reduction, kernel_size=1, padding=0)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(channels // reduction, channels, kernel_size=1, padding=0)
        self.sigmoid = nn.Sigmoid()
        _init(self)


class SpatialAttentionBlock(nn.Module):
    """sigmoid(phi(relu(bn(down(x))))) -> one attention map shared by all channels."""

    def __init__(self, in_features, attn_features, up_factor, normalize_attn=False):
        super().__init__()
        self.up_factor = up_factor
        self.normalize_attn = normalize_attn
        self.down = nn.Conv2d(in_features, attn_features, kernel_size=1, padding=0, bias=False)
        self.phi = nn.Conv2d(attn_features, out_channels=1, kernel_size=1, padding=0, bias=True)
        self.relu = nn.ReLU(inplace=True)
        self.bn = nn.BatchNorm2d(attn_features)
        _init(self)


class _MRF(nn.Module):
    """cat([skip, relu(bn(convT4x4s2(lo)))], dim=1)."""

    def __init__(self, inchannels):
        super().__init__()
        self.up = nn.Sequential(
            nn.ConvTranspose2d(inchannels[0], inchannels[0], kernel_size=4, stride=2, padding=1),
            nn.BatchNorm2d(inchannels[0]),
            nn.ReLU(inplace=True))
        _init(self)


class DualAttBlock(nn.Module):
    def __init__(self, inchannels=[128, 256], outchannels=256):
        super().__init__()
        inchs = sum(inchannels)
        self.mrf = _MRF(inchannels)
        self.spatialAttn = SpatialAttentionBlock(outchannels, int(outchannels / 4), 2)
        self.channelAttn = SEModule(outchannels, 16)
        self.c3x3rb = nn.Sequential(nn.Conv2d(inchs, outchannels, kernel_size=3, padding=1),
                                    nn.BatchNorm2d(outchannels),
                                    nn.ReLU(inplace=True))
        _init(self)

    def _body(self, tp, lo, skip, mcat=None):
        return dual_att_body(tp, self, lo, skip, mcat)

    def forward(self, x):
        """x = [low-res feature, skip feature] -> (out, spatial attention map)."""
        if len(x) != 2:
            raise NotImplementedError("DualAttBlock: the CUDA path implements the two-input form SAUNet uses")
        out, spatial = engine.run(self, lambda tp, lo, skip: list(self._body(tp, lo, skip)), [x[0], x[1]])
        return out, spatial
