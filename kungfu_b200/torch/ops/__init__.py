"""`kungfu.torch.ops` parity (srcs/python/kungfu/torch/ops/collective.py:8-55)."""
from __future__ import annotations

import torch

from ...ops import collective as _c
from ...ops.collective import Handle, all_gather, barrier, consensus, wait_all_handles, wait_handle


def all_reduce_fn(x: torch.Tensor, op=None) -> torch.Tensor:
    """Out-of-place all-reduce (default op: sum, as in the reference)."""
    return _c.all_reduce(x, op or "sum")


def inplace_all_reduce_op(x: torch.Tensor, op=None) -> None:
    """Synchronous in-place all-reduce (reference: torch/ops/collective.py:8-27)."""
    _c.all_reduce_(x, op or "sum")


def inplace_all_reduce_async_op(x: torch.Tensor, name: str = "", op=None) -> Handle:
    """Asynchronous in-place all-reduce; returns a handle for `wait_handle` / `wait_all_handles` (reference: :29-48)."""
    return _c.all_reduce_async_(x, name, op or "sum")


def inplace_broadcast_op(x: torch.Tensor, root: int = 0) -> None:
    """In-place broadcast from `root` (reference: torch/ops/collective.py broadcast)."""
    _c.broadcast_(x, root)


def inplace_broadcast_async_op(x: torch.Tensor, name: str = "", root: int = 0) -> Handle:
    """Asynchronous in-place broadcast from `root`; returns a handle (reference: torch/ops/collective.py:29-31)."""
    return _c.broadcast_async_(x, name, root)


def broadcast_parameters(state_dict, root: int = 0) -> None:
    """Broadcast every tensor of `state_dict` from `root` (reference: torch/ops/collective.py:42-47)."""
    for name, value in state_dict.items():
        if isinstance(value, torch.Tensor):
            _c.broadcast_(value, root, name)


__all__ = ["all_reduce_fn", "broadcast_parameters", "inplace_all_reduce_async_op", "inplace_all_reduce_op",
           "inplace_broadcast_op", "inplace_broadcast_async_op", "wait_handle", "wait_all_handles", "all_gather", "barrier", "consensus", "Handle"]
