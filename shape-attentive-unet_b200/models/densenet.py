"""DenseNet-121 feature extractor: the parameter tree of torchvision.models.densenet121 (densenet.py:31-133,
160-210 in torchvision 0.26), which the reference uses as the SAUNet encoder (models/models.py:271,304-313).

torchvision is a third-party dependency of the reference and is NOT imported here; this file only rebuilds the
module/parameter naming (so reference checkpoints load key-for-key) and torchvision's initialisation.  The
arithmetic is in saunet_b200.blocks.dense_block_body / transition_body.
"""
from collections import OrderedDict

import torch.nn as nn


class _DenseLayer(nn.Module):
    def __init__(self, num_input_features, growth_rate, bn_size):
        super().__init__()
        self.norm1 = nn.BatchNorm2d(num_input_features)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv1 = nn.Conv2d(num_input_features, bn_size * growth_rate, kernel_size=1, stride=1, bias=False)
        self.norm2 = nn.BatchNorm2d(bn_size * growth_rate)
        self.relu2 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(bn_size * growth_rate, growth_rate, kernel_size=3, stride=1, padding=1, bias=False)


class _DenseBlock(nn.ModuleDict):
    def __init__(self, num_layers, num_input_features, bn_size, growth_rate):
        super().__init__()
        for i in range(num_layers):
            self.add_module("denselayer%d" % (i + 1),
                            _DenseLayer(num_input_features + i * growth_rate, growth_rate, bn_size))
        self.out_features = num_input_features + num_layers * growth_rate


class _Transition(nn.Sequential):
    def __init__(self, num_input_features, num_output_features):
        super().__init__()
        self.norm = nn.BatchNorm2d(num_input_features)
        self.relu = nn.ReLU(inplace=True)
        self.conv = nn.Conv2d(num_input_features, num_output_features, kernel_size=1, stride=1, bias=False)
        self.pool = nn.AvgPool2d(kernel_size=2, stride=2)


class DenseNet(nn.Module):
    def __init__(self, growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4,
                 num_classes=1000):
        super().__init__()
        self.features = nn.Sequential(OrderedDict([
            ("conv0", nn.Conv2d(3, num_init_features, kernel_size=7, stride=2, padding=3, bias=False)),
            ("norm0", nn.BatchNorm2d(num_init_features)),
            ("relu0", nn.ReLU(inplace=True)),
            ("pool0", nn.MaxPool2d(kernel_size=3, stride=2, padding=1)),
        ]))
        nf = num_init_features
        for i, num_layers in enumerate(block_config):
            block = _DenseBlock(num_layers, nf, bn_size, growth_rate)
            self.features.add_module("denseblock%d" % (i + 1), block)
            nf = nf + num_layers * growth_rate
            if i != len(block_config) - 1:
                self.features.add_module("transition%d" % (i + 1), _Transition(nf, nf // 2))
                nf = nf // 2
        self.features.add_module("norm5", nn.BatchNorm2d(nf))
        # never used by SAUNet, kept so the state_dict / parameter set matches the reference (1 025 000 params)
        self.classifier = nn.Linear(nf, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.constant_(m.bias, 0)


def densenet121(pretrained=False):
    if pretrained:
        raise RuntimeError("densenet121(pretrained=True) needs the ImageNet checkpoint download, which this offline "
                           "build cannot do; load weights with SAUNet.load_state_dict / ModelBuilder(weights=...)")
    return DenseNet(32, (6, 12, 24, 16), 64)
