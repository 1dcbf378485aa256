"""Adaptation ops: resize, step-based schedules, set_tree.

Reference: srcs/python/kungfu/tensorflow/ops/adapt.py:5-62 and the C++ kernels `KungfuResizeCluster(FromURL)`
(srcs/cpp/src/tensorflow/ops/cpu/control.cpp), `KungfuStepBasedSchedule` (cpu/elastic.cpp:16-79), `KungfuSetTree`
(cpu/adaptation.cpp).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from ..python import resize as _resize
from ..runtime import get_runtime


def resize(n: Optional[int] = None) -> Tuple[bool, bool]:
    """Resize the cluster to `n` workers -> (changed, detached)."""
    return _resize(n)


resize_cluster = resize


def resize_cluster_from_url() -> Tuple[bool, bool]:
    """Adopt whatever cluster the config server currently holds -> (changed, detached)."""
    return _resize(None)


class StepBasedSchedule:
    """Piece-wise constant cluster size by step. Grammar: "size:steps,size:steps,..." (e.g. "1:3,2:3,4:3").

    Out-of-range steps give `default` unless `strict` (then ValueError) — cpu/elastic.cpp:39-77."""

    def __init__(self, config: str, default: int = 1, strict: bool = False) -> None:
        if not config:
            raise ValueError("config can't be empty")
        self.ranges: List[Tuple[int, int, int]] = []
        off = 0
        for part in config.split(","):
            kv = part.split(":")
            if len(kv) != 2:
                raise ValueError(f"invalid schedule config {config!r}")
            size, steps = int(kv[0]), int(kv[1])
            self.ranges.append((off, off + steps, size))
            off += steps
        self.total_steps = off
        self.default = default
        self.strict = strict

    def __call__(self, step: int) -> int:
        step = int(step)
        for lo, hi, size in self.ranges:
            if lo <= step < hi:
                return size
        if self.strict:
            raise ValueError(f"schedule not found for step {step}")
        return self.default


def step_based_schedule(config: str, step: int, default: int = 1, strict: bool = False) -> int:
    return StepBasedSchedule(config, default, strict)(step)


def set_tree(tree: Sequence[int]) -> None:
    """Switch the CPU all-reduce strategy to the rooted tree given as a father[] array (father[root] == root).
    All peers must pass the same tree (checked by consensus). On the GPU plane topology graphs are moot (uniform
    NVSwitch): the call only re-roots host-side collectives."""
    rt = get_runtime()
    f = getattr(rt.control, "set_tree", None)
    if f is None:
        if rt.size == 1:
            return
        raise RuntimeError("set_tree needs the native runtime")
    if isinstance(tree, torch.Tensor):
        tree = tree.tolist()
    f([int(x) for x in tree])
