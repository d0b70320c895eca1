"""GatedSpatialConv2d (models/GSConv.py:16-62 of the reference): the gated shape-stream convolution.

alphas = sigmoid(BN1(conv1x1_{C+1->1}(relu(conv1x1_{C+1->C+1}(BN_{C+1}(cat[x, g]))))))
out    = conv1x1(x * (alphas + 1), weight)           -> (out, alphas)
"""
import torch.nn as nn
from torch.nn.modules.conv import _ConvNd
from torch.nn.modules.utils import _pair

from saunet_b200 import engine
from saunet_b200.blocks import gsconv_body
from . import norm as mynn


class GatedSpatialConv2d(_ConvNd):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, groups=1, bias=False):
        super().__init__(in_channels, out_channels, _pair(kernel_size), _pair(stride), _pair(padding), _pair(dilation),
                         False, _pair(0), groups, bias, "zeros")
        self._gate_conv = nn.Sequential(
            mynn.Norm2d(in_channels + 1),
            nn.Conv2d(in_channels + 1, in_channels + 1, 1),
            nn.ReLU(),
            nn.Conv2d(in_channels + 1, 1, 1),
            mynn.Norm2d(1),
            nn.Sigmoid(),
        )

    def _body(self, tp, x, g):
        return gsconv_body(tp, self, x, g)

    def forward(self, input_features, gating_features):
        out, alphas = engine.run(self, lambda tp, x, g: list(self._body(tp, x, g)), [input_features, gating_features])
        return out, alphas

    def reset_parameters(self):
        nn.init.xavier_normal_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)
