"""VGG (configurations A / D / E = VGG11 / 16 / 19): the second model of the reference's scalability charts.

Reference: benchmarks/system/benchmark_kungfu.py:58-66 takes `tf.keras.applications.VGG16` (no batch norm, 138,357,544
parameters in 32 tensors, one of them the 102.8 M-element first classifier matrix — the message shape that makes VGG the
communication-heavy case; `model_sizes.py:1-5`). Same architecture here, channels_last / bf16 friendly: the convolutions run in
cuDNN, every bias+ReLU pair is one in-place pass, the classifier matrices are plain cuBLAS GEMMs.
"""
from __future__ import annotations

from typing import List, Union

import torch
import torch.nn as nn

CFG = {
    "A": [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
    "D": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
    "E": [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"],
}


class VGG(nn.Module):
    def __init__(self, cfg: List[Union[int, str]], num_classes: int = 1000, dropout: float = 0.5, hidden: int = 4096) -> None:
        super().__init__()
        layers: List[nn.Module] = []
        c = 3
        for v in cfg:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(c, int(v), 3, padding=1), nn.ReLU(inplace=True)]
                c = int(v)
        self.features = nn.Sequential(*layers)
        self.avgpool = nn.AdaptiveAvgPool2d((7, 7))
        self.classifier = nn.Sequential(nn.Linear(c * 49, hidden), nn.ReLU(inplace=True), nn.Dropout(dropout), nn.Linear(hidden, hidden),
                                        nn.ReLU(inplace=True), nn.Dropout(dropout), nn.Linear(hidden, num_classes))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.zeros_(m.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.avgpool(self.features(x))
        # flatten in NCHW order whatever the memory format, so that checkpoints are layout independent
        return self.classifier(torch.flatten(x.contiguous(), 1))


def vgg11(num_classes: int = 1000) -> VGG:
    return VGG(CFG["A"], num_classes)


def vgg16(num_classes: int = 1000) -> VGG:
    return VGG(CFG["D"], num_classes)


def vgg19(num_classes: int = 1000) -> VGG:
    return VGG(CFG["E"], num_classes)
