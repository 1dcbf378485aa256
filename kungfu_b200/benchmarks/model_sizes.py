"""Gradient tensor shapes of the models used by the communication micro-benchmarks.

Reference: srcs/python/kungfu/tensorflow/v1/benchmarks/model_sizes.py:1-81 (ResNet-50: 214 tensors / 25.58 M elements as
built by tf.keras; VGG16: 32 tensors / 138.36 M; "bert": 599 tensors). Ours are derived from the actual torch modules
(torchvision-identical ResNet-50: 161 tensors / 25.56 M; BERT-large from `models.bert`), VGG16 sizes are listed.
"""
from __future__ import annotations

from typing import List

VGG16 = [1728, 64, 36864, 64, 73728, 128, 147456, 128, 294912, 256, 589824, 256, 589824, 256, 1179648, 512, 2359296, 512, 2359296, 512,
         2359296, 512, 2359296, 512, 2359296, 512, 102760448, 4096, 16777216, 4096, 4096000, 1000]


def sizes(model: str) -> List[int]:
    model = model.lower()
    if model == "vgg16":
        return list(VGG16)
    if model == "slp":
        return [7840, 10]
    if model in ("resnet50", "resnet18", "resnet101", "vgg11", "vgg19", "inception_v3"):
        from .. import models

        return [p.numel() for p in models.build(model).parameters()]
    if model in ("bert", "bert-large", "bert_large"):
        from ..models import bert

        return bert.parameter_sizes(bert.BertConfig.large())
    if model in ("bert-base", "bert_base"):
        from ..models import bert

        return bert.parameter_sizes(bert.BertConfig.base())
    raise ValueError(f"unknown model {model!r}")
