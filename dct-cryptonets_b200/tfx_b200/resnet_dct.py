"""Circuit topology source: the DCT-CryptoNets ResNet-20 / ResNet-18 feature extractors.

Own restatement of the topology defined by the reference at models/backbone.py:18-58 (SimpleBlock),
:107-184 (ResNetDCT), :291-342 (factories) and the stem table :347-582 — the part of the reference that defines
the work of the encrypted path (which convs, where the table lookups sit).  It exists because /root/reference is
not present on the GPU box; tests/test_circuit_cpu.py::test_topology_matches_reference_modules checks it against the reference modules in the
build container.  Attribute names (.trunk, .final_feat_dim, C1/BN1/relu1/C2/BN2/relu2/shortcut/BNshortcut)
follow the reference so the same compile front-end accepts either.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

# stem variants keyed like the reference table: "<first block width>_<input channels>_<input size>"
# (conv1 kernel, stride, padding, stem relu, maxpool kernel/stride or None, final avgpool kernel)
STEMS = {
    "16_3_32": (3, 1, 1, True, None, 7),
    "48_3_32": (3, 1, 1, True, None, 7),
    "48_24_32": (1, 1, 0, True, None, 16),
    "48_24_64": (1, 1, 0, True, None, 32),
    "48_24_8": (1, 1, 0, True, None, 3),
    "48_24_16": (1, 1, 0, True, None, 7),
    "48_48_8": (1, 1, 0, True, None, 3),
    "48_48_16": (1, 1, 0, True, None, 7),
    "64_48_16": (1, 1, 0, True, None, 3),
    # deviation: the reference has no '64_24_16' entry; 1x1 stem, no pool, avgpool over the 2x2 map
    "64_24_16": (1, 1, 0, True, None, 2),
    "64_6_32": (1, 1, 0, False, None, 3),
    "64_3_32": (3, 1, 1, True, None, 3),
    "64_6_56": (1, 1, 0, False, None, 5),
    "64_12_56": (1, 1, 0, False, None, 5),
    "64_24_56": (1, 1, 0, False, None, 5),
    "64_48_56": (1, 1, 0, False, None, 5),
    "64_64_56": (1, 1, 0, False, None, 5),
    "64_192_56": (1, 1, 0, False, None, 5),
    # RGB stems with a max-pool (reference models/backbone.py:447-481): 7x7 stride-2 conv, ReLU, MaxPool2d(3, 2, padding 1)
    "64_3_128": (7, 2, 3, True, (3, 2), 3),
    "64_3_224": (7, 2, 3, True, (3, 2), 7),
    "64_3_448": (7, 2, 3, True, (3, 2), 14),
}


def _init(m: nn.Module):
    if isinstance(m, nn.Conv2d):
        fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
        m.weight.data.normal_(0.0, math.sqrt(2.0 / fan))
    elif isinstance(m, nn.BatchNorm2d):
        m.weight.data.fill_(1.0)
        m.bias.data.zero_()


class ResidualBlock(nn.Module):
    """conv3x3-BN-ReLU-conv3x3-BN (+ identity or 1x1-conv-BN shortcut) -ReLU"""

    def __init__(self, cin: int, cout: int, downsample: bool):
        super().__init__()
        s = 2 if downsample else 1
        self.C1 = nn.Conv2d(cin, cout, 3, stride=s, padding=1, bias=False)
        self.BN1 = nn.BatchNorm2d(cout)
        self.relu1 = nn.ReLU()
        self.C2 = nn.Conv2d(cout, cout, 3, padding=1, bias=False)
        self.BN2 = nn.BatchNorm2d(cout)
        self.relu2 = nn.ReLU()
        self.shortcut_type = "identity" if cin == cout else "1x1"
        if cin != cout:
            self.shortcut = nn.Conv2d(cin, cout, 1, stride=s, bias=False)
            self.BNshortcut = nn.BatchNorm2d(cout)
        for m in self.children():
            _init(m)

    def forward(self, x):
        y = self.BN2(self.C2(self.relu1(self.BN1(self.C1(x)))))
        sc = x if self.shortcut_type == "identity" else self.BNshortcut(self.shortcut(x))
        return self.relu2(y + sc)


class ResNetDCTFeatures(nn.Module):
    def __init__(self, blocks_per_stage, widths, in_channels: int, img_size: int, skip_single_downsample: bool):
        super().__init__()
        key = f"{widths[0]}_{in_channels}_{img_size}"
        if key not in STEMS:
            raise KeyError(f"no stem variant '{key}'")
        k, s, p, stem_relu, pool, avg = STEMS[key]
        layers = [nn.Conv2d(in_channels, widths[0], k, stride=s, padding=p, bias=False), nn.BatchNorm2d(widths[0])]
        _init(layers[0]); _init(layers[1])
        if stem_relu:
            layers.append(nn.ReLU())
        if pool is not None:
            layers.append(nn.MaxPool2d(pool[0], stride=pool[1], padding=1))
        cin = widths[0]
        for stage, (nb, cout) in enumerate(zip(blocks_per_stage, widths)):
            for j in range(nb):
                first_down = 2 if skip_single_downsample else 1
                layers.append(ResidualBlock(cin, cout, downsample=(stage >= first_down and j == 0)))
                cin = cout
        layers += [nn.AvgPool2d(avg), nn.Flatten()]
        self.trunk = nn.Sequential(*layers)
        self.final_feat_dim = cin

    def forward(self, x):
        return self.trunk(x)


def resnet20_dct(in_channels: int = 24, img_size: int = 16, skip_single_downsample: bool = True) -> ResNetDCTFeatures:
    """DCT-CryptoNets ResNet-20: 3 stages x 3 blocks, widths 48/56/64 (reference backbone.py:291-302)."""
    return ResNetDCTFeatures([3, 3, 3], [48, 56, 64], in_channels, img_size, skip_single_downsample)


def resnet18_dct(in_channels: int = 24, img_size: int = 16) -> ResNetDCTFeatures:
    """ResNet-18: 4 stages x 2 blocks, widths 64/128/256/512 (reference backbone.py:320-329)."""
    return ResNetDCTFeatures([2, 2, 2, 2], [64, 128, 256, 512], in_channels, img_size, False)


class FeatureClassifier(nn.Module):
    """feature extractor + clear linear classifier, the shape of the reference's BaselineTrain (utils.py:14-47):
    only .feature is compiled to FHE; .classifier runs in the clear on decrypted features."""

    def __init__(self, feature: nn.Module, num_class: int = 10):
        super().__init__()
        self.feature = feature
        self.classifier = nn.Linear(feature.final_feat_dim, num_class)
        self.classifier.bias.data.zero_()

    def forward(self, x):
        f = self.feature(x)
        return f, self.classifier(f)
