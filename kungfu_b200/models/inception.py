"""Inception v3: the third model of the reference's scalability charts.

Reference: benchmarks/system/benchmark_kungfu.py:58-66 takes `tf.keras.applications.InceptionV3` (299x299 input, ~23.9 M
parameters without the auxiliary head, every convolution followed by BatchNorm + ReLU). Every conv unit below is
`conv -> BNAct` with the fused sm_100a BN+ReLU kernels of `ops/bn_act.py` where the channel count allows it (power-of-two
multiples of 8) and the PyTorch path otherwise; concatenations stay in channels_last so that they are plain row copies.
"""
from __future__ import annotations

from typing import Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .resnet import BNAct

_size = Union[int, Tuple[int, int]]


class ConvBN(nn.Module):
    def __init__(self, i: int, o: int, kernel: _size, stride: _size = 1, padding: _size = 0) -> None:
        super().__init__()
        self.conv = nn.Conv2d(i, o, kernel, stride, padding, bias=False)
        self.bn = BNAct(o, relu=True)
        self.bn.eps = 1e-3

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.bn(self.conv(x))


class InceptionA(nn.Module):  # 35 x 35
    def __init__(self, i: int, pool: int) -> None:
        super().__init__()
        self.b1 = ConvBN(i, 64, 1)
        self.b5 = nn.Sequential(ConvBN(i, 48, 1), ConvBN(48, 64, 5, padding=2))
        self.b3 = nn.Sequential(ConvBN(i, 64, 1), ConvBN(64, 96, 3, padding=1), ConvBN(96, 96, 3, padding=1))
        self.bp = ConvBN(i, pool, 1)

    def forward(self, x):
        return torch.cat([self.b1(x), self.b5(x), self.b3(x), self.bp(F.avg_pool2d(x, 3, 1, 1))], 1)


class InceptionB(nn.Module):  # 35 -> 17
    def __init__(self, i: int) -> None:
        super().__init__()
        self.b3 = ConvBN(i, 384, 3, stride=2)
        self.bd = nn.Sequential(ConvBN(i, 64, 1), ConvBN(64, 96, 3, padding=1), ConvBN(96, 96, 3, stride=2))

    def forward(self, x):
        return torch.cat([self.b3(x), self.bd(x), F.max_pool2d(x, 3, 2)], 1)


class InceptionC(nn.Module):  # 17 x 17, factorised 7x7
    def __init__(self, i: int, c7: int) -> None:
        super().__init__()
        self.b1 = ConvBN(i, 192, 1)
        self.b7 = nn.Sequential(ConvBN(i, c7, 1), ConvBN(c7, c7, (1, 7), padding=(0, 3)), ConvBN(c7, 192, (7, 1), padding=(3, 0)))
        self.bd = nn.Sequential(ConvBN(i, c7, 1), ConvBN(c7, c7, (7, 1), padding=(3, 0)), ConvBN(c7, c7, (1, 7), padding=(0, 3)),
                                ConvBN(c7, c7, (7, 1), padding=(3, 0)), ConvBN(c7, 192, (1, 7), padding=(0, 3)))
        self.bp = ConvBN(i, 192, 1)

    def forward(self, x):
        return torch.cat([self.b1(x), self.b7(x), self.bd(x), self.bp(F.avg_pool2d(x, 3, 1, 1))], 1)


class InceptionD(nn.Module):  # 17 -> 8
    def __init__(self, i: int) -> None:
        super().__init__()
        self.b3 = nn.Sequential(ConvBN(i, 192, 1), ConvBN(192, 320, 3, stride=2))
        self.b7 = nn.Sequential(ConvBN(i, 192, 1), ConvBN(192, 192, (1, 7), padding=(0, 3)), ConvBN(192, 192, (7, 1), padding=(3, 0)),
                                ConvBN(192, 192, 3, stride=2))

    def forward(self, x):
        return torch.cat([self.b3(x), self.b7(x), F.max_pool2d(x, 3, 2)], 1)


class InceptionE(nn.Module):  # 8 x 8, expanded filter bank
    def __init__(self, i: int) -> None:
        super().__init__()
        self.b1 = ConvBN(i, 320, 1)
        self.b3 = ConvBN(i, 384, 1)
        self.b3a, self.b3b = ConvBN(384, 384, (1, 3), padding=(0, 1)), ConvBN(384, 384, (3, 1), padding=(1, 0))
        self.bd = nn.Sequential(ConvBN(i, 448, 1), ConvBN(448, 384, 3, padding=1))
        self.bda, self.bdb = ConvBN(384, 384, (1, 3), padding=(0, 1)), ConvBN(384, 384, (3, 1), padding=(1, 0))
        self.bp = ConvBN(i, 192, 1)

    def forward(self, x):
        b3 = self.b3(x)
        bd = self.bd(x)
        return torch.cat([self.b1(x), self.b3a(b3), self.b3b(b3), self.bda(bd), self.bdb(bd), self.bp(F.avg_pool2d(x, 3, 1, 1))], 1)


class InceptionAux(nn.Module):
    def __init__(self, i: int, num_classes: int) -> None:
        super().__init__()
        self.conv0, self.conv1 = ConvBN(i, 128, 1), ConvBN(128, 768, 5)
        self.fc = nn.Linear(768, num_classes)

    def forward(self, x):
        x = self.conv1(self.conv0(F.avg_pool2d(x, 5, 3)))
        return self.fc(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1))


class InceptionV3(nn.Module):
    """`aux_logits=False` (default) is the keras-applications graph the reference benchmarks; with `aux_logits=True` training
    mode returns (logits, aux_logits) like torchvision."""

    def __init__(self, num_classes: int = 1000, aux_logits: bool = False, dropout: float = 0.5) -> None:
        super().__init__()
        self.stem = nn.Sequential(ConvBN(3, 32, 3, stride=2), ConvBN(32, 32, 3), ConvBN(32, 64, 3, padding=1), nn.MaxPool2d(3, 2),
                                  ConvBN(64, 80, 1), ConvBN(80, 192, 3), nn.MaxPool2d(3, 2))
        self.mixed5 = nn.Sequential(InceptionA(192, 32), InceptionA(256, 64), InceptionA(288, 64))
        self.mixed6 = nn.Sequential(InceptionB(288), InceptionC(768, 128), InceptionC(768, 160), InceptionC(768, 160), InceptionC(768, 192))
        self.aux = InceptionAux(768, num_classes) if aux_logits else None
        self.mixed7 = nn.Sequential(InceptionD(768), InceptionE(1280), InceptionE(2048))
        self.dropout = nn.Dropout(dropout)
        self.fc = nn.Linear(2048, num_classes)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.trunc_normal_(m.weight, std=0.1 if m is not self.fc else 0.01, a=-2, b=2)

    def forward(self, x: torch.Tensor):
        x = self.mixed6(self.mixed5(self.stem(x)))
        aux = self.aux(x) if (self.aux is not None and self.training) else None
        x = self.mixed7(x)
        x = self.fc(self.dropout(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1)))
        return (x, aux) if aux is not None else x


def inception_v3(num_classes: int = 1000, aux_logits: bool = False) -> InceptionV3:
    return InceptionV3(num_classes, aux_logits)
