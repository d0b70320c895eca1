"""BasicBlock (models/resnet.py:30-59 of the reference), the shape-stream residual block of SAUNet.

Only BasicBlock is on the hot path; the ResNet trunk builders of the reference file
are unused by SAUNet and not rebuilt.
"""
import torch.nn as nn

from lib.nn import SynchronizedBatchNorm2d
from saunet_b200 import engine
from saunet_b200.blocks import basic_block_body


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = SynchronizedBatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = SynchronizedBatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def _body(self, tp, x):
        return basic_block_body(tp, self, x)

    def forward(self, x):
        return engine.run(self, lambda tp, xb: [self._body(tp, xb)], [x])[0]
