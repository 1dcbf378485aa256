"""ResNet family (ImageNet), architecture-identical to torchvision's `resnet50` (25,557,032 parameters, 161
trainable tensors) — the model of the reference's headline benchmark (benchmarks/system/benchmark_kungfu_torch.py:58
`getattr(models, args.model)()`; README.md:201-219).

B200 specifics: built for channels_last + bf16 from the start; BatchNorm+ReLU(+residual add) are expressed as one
module (`BNAct`) so that a fused sm_100a kernel can replace the three memory-bound element-wise passes.
"""
from __future__ import annotations

from typing import List, Optional, Type

import torch
import torch.nn as nn
import torch.nn.functional as F


class BNAct(nn.BatchNorm2d):
    """BatchNorm2d followed by an optional residual add and ReLU."""

    def __init__(self, ch: int, relu: bool = True, zero_init: bool = False) -> None:
        super().__init__(ch)
        self.relu = relu
        if zero_init:
            nn.init.zeros_(self.weight)

    fused = True  # use the fused sm_100a kernel (ops/bn_act.py) when the layout/dtype allows it
    external_counter = False  # the owning model bumps every `num_batches_tracked` with one multi-tensor launch

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:  # type: ignore[override]
        if BNAct.fused and x.is_cuda:
            from ..ops.bn_act import bn_act

            if self.training and self.track_running_stats and self.num_batches_tracked is not None and not self.external_counter:
                self.num_batches_tracked.add_(1)
            return bn_act(x, self.weight, self.bias, self.running_mean, self.running_var, residual, self.relu,
                          self.training or not self.track_running_stats, self.momentum if self.momentum is not None else 0.1, self.eps)
        y = super().forward(x)
        if residual is not None:
            y = y + residual
        return F.relu(y, inplace=True) if self.relu else y


def conv3x3(i: int, o: int, stride: int = 1) -> nn.Conv2d:
    return nn.Conv2d(i, o, 3, stride, 1, bias=False)


def conv1x1(i: int, o: int, stride: int = 1) -> nn.Conv2d:
    return nn.Conv2d(i, o, 1, stride, 0, bias=False)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: Optional[nn.Module] = None) -> None:
        super().__init__()
        self.conv1 = conv1x1(inplanes, planes)
        self.bn1 = BNAct(planes)
        self.conv2 = conv3x3(planes, planes, stride)
        self.bn2 = BNAct(planes)
        self.conv3 = conv1x1(planes, planes * 4)
        self.bn3 = BNAct(planes * 4)
        self.downsample = downsample

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        identity = x if self.downsample is None else self.downsample(x)
        out = self.bn1(self.conv1(x))
        out = self.bn2(self.conv2(out))
        return self.bn3(self.conv3(out), identity)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: Optional[nn.Module] = None) -> None:
        super().__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = BNAct(planes)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = BNAct(planes)
        self.downsample = downsample

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        identity = x if self.downsample is None else self.downsample(x)
        out = self.bn1(self.conv1(x))
        return self.bn2(self.conv2(out), identity)


class _Down(nn.Sequential):
    def __init__(self, i: int, o: int, stride: int) -> None:
        super().__init__(conv1x1(i, o, stride), BNAct(o, relu=False))


def stem_conv_s2d(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """The 7x7 / stride-2 / pad-3 stem convolution evaluated as a 4x4 / stride-1 convolution on a space-to-depth(2) input.

    Identical math (the 7x7 kernel is zero-padded to 8x8 at the front and folded 2x2 into channels), but the GEMM sees
    C_in = 12 (padded to 16) instead of 3, so cuDNN runs its Blackwell tensor-core kernels instead of the legacy
    `sm80_xmma` / padded `align8` paths that make the C_in = 3 stem cost 2.2 ms per step at batch 256
    (profiles/resnet50_step_cupti.md). The parameter keeps its [64, 3, 7, 7] shape; gradients flow back through the fold.
    """
    N, C, H, W = x.shape
    O = weight.shape[0]
    w8 = F.pad(weight, (1, 0, 1, 0))  # [O, C, 8, 8], zero tap in front: kh' = kh + 1 = 2a + b
    w4 = w8.reshape(O, C, 4, 2, 4, 2).permute(0, 3, 5, 1, 2, 4).reshape(O, 4 * C, 4, 4)  # channels ordered (b, b2, c)
    cpad = (-4 * C) % 8
    w4 = F.pad(w4, (0, 0, 0, 0, 0, cpad)).contiguous(memory_format=torch.channels_last)
    t = x.permute(0, 2, 3, 1)  # NHWC view
    t = t.reshape(N, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(N, H // 2, W // 2, 4 * C)
    t = F.pad(t, (0, cpad, 2, 1, 2, 1))  # channels -> multiple of 8; ih = 2 (oh - 2 + a) + b => pad 2 before, 1 after
    return F.conv2d(t.permute(0, 3, 1, 2), w4)


class ResNet(nn.Module):
    s2d_stem = True  # use `stem_conv_s2d` for CUDA inputs with even spatial size

    def __init__(self, block: Type[nn.Module], layers: List[int], num_classes: int = 1000) -> None:
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = BNAct(64)
        from ..ops.pool import MaxPool3x3s2

        self.maxpool = MaxPool3x3s2()  # fused sm_100a kernel on channels_last bf16, F.max_pool2d otherwise
        self.layer1 = self._make_layer(block, 64, layers[0], 1)
        self.layer2 = self._make_layer(block, 128, layers[1], 2)
        self.layer3 = self._make_layer(block, 256, layers[2], 2)
        self.layer4 = self._make_layer(block, 512, layers[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def _make_layer(self, block, planes: int, blocks: int, stride: int) -> nn.Sequential:
        down = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            down = _Down(self.inplanes, planes * block.expansion, stride)
        layers = [block(self.inplanes, planes, stride, down)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training and BNAct.fused and x.is_cuda:
            bns = [m for m in self.modules() if isinstance(m, BNAct)]
            for m in bns:
                m.external_counter = True
            torch._foreach_add_([m.num_batches_tracked for m in bns], 1)  # 1 launch instead of 53
        if ResNet.s2d_stem and x.is_cuda and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
            x = stem_conv_s2d(x, self.conv1.weight)
        else:
            x = self.conv1(x)
        x = self.maxpool(self.bn1(x))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(self.avgpool(x), 1))


def resnet18(num_classes: int = 1000) -> ResNet:
    return ResNet(BasicBlock, [2, 2, 2, 2], num_classes)


def resnet50(num_classes: int = 1000) -> ResNet:
    return ResNet(Bottleneck, [3, 4, 6, 3], num_classes)


def resnet101(num_classes: int = 1000) -> ResNet:
    return ResNet(Bottleneck, [3, 4, 23, 3], num_classes)


def to_b200(model: nn.Module, dtype: torch.dtype = torch.bfloat16, device="cuda") -> nn.Module:
    """channels_last + low-precision weights; the fused optimizers keep the fp32 master copy."""
    return model.to(device=device, dtype=dtype, memory_format=torch.channels_last)
