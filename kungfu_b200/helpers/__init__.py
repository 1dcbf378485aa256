"""Dataset helpers: IDX codec, MNIST / CIFAR loaders, an ImageNet directory reader and synthetic inputs.

Reference: srcs/python/kungfu/tensorflow/v1/helpers/{__init__,idx,mnist,cifar,imagenet,utils}.py. Arrays come back as numpy
(host side, so they can sit in pinned memory); `to_device_batches` turns a loaded set into the uint8-over-PCIe pipeline of
`kungfu_b200.trainer` (normalisation happens on the GPU, the host ships bytes).
"""
from __future__ import annotations

from . import cifar, idx, imagenet, mnist  # noqa: F401
from .utils import show_rate, show_size  # noqa: F401


def random_input(sample_shape, feature_categories, device="cpu", dtype=None, generator=None):
    """One synthetic (samples, label) pair: uniform samples of `sample_shape`, one label in [0, feature_categories - 1)
    (helpers/__init__.py:4-10 — the reference's label range excludes the last class too)."""
    import torch

    samples = torch.rand(tuple(sample_shape), device=device, dtype=dtype or torch.float32, generator=generator)
    labels = torch.randint(0, max(1, feature_categories - 1), (), device=device, generator=generator, dtype=torch.int32)
    return samples, labels
