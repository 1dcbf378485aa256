"""`kungfu.torch.ops.clib` — dispatch tables from tensor type to the native operator.

Reference: srcs/python/kungfu/torch/ops/clib.py:4-40 imports the compiled `kungfu_torch_ops` extension and builds
`all_reduce_op_map` / `all_gather_op_map` / `all_reduce_async_op_map` / `broadcast_async_op_map` keyed by `tensor.type()`
('torch.FloatTensor', 'torch.cuda.FloatTensor'), float32 only. Here the native side is `_kfrt` (host plane) and `_kfcuda` (sm_100a
collective kernels over peer-mapped HBM); the tables cover every dtype both planes reduce, and the entries keep the reference's call
shape `op(x, y, type, op_name[, name])` so that code written against the tables keeps working.
"""
from __future__ import annotations

import torch

from ...ops import collective as _c

_DTYPES = ("Float", "Double", "Half", "BFloat16", "Int", "Long", "Short", "Char", "Byte")
_HOST = tuple(f"torch.{d}Tensor" for d in _DTYPES)
_CUDA = tuple(f"torch.cuda.{d}Tensor" for d in _DTYPES)


def _all_reduce(x, y, _type=None, op="sum", name=""):
    if y is not x:
        y.copy_(x)
    _c.all_reduce_(y, op, name)


def _all_reduce_async(x, y, _type=None, op="sum", name=""):
    if y is not x:
        y.copy_(x)
    return _c.all_reduce_async_(y, name, op)


def _broadcast_async(x, y, _type=None, name=""):
    if y is not x:
        y.copy_(x)
    return _c.broadcast_async_(y, name)


def _all_gather(x, y, _type=None):
    y.copy_(_c.all_gather(x).view(y.shape))


class _Ops:
    """Stands where the reference has the extension module: the entry points by their native names."""

    all_reduce_cpu = staticmethod(_all_reduce)
    all_reduce_cuda = staticmethod(_all_reduce)
    all_reduce_cuda_async = staticmethod(_all_reduce_async)
    broadcast_cuda_async = staticmethod(_broadcast_async)
    all_gather_cpu = staticmethod(_all_gather)
    all_gather_cuda = staticmethod(_all_gather)
    wait_handle = staticmethod(_c.wait_handle)
    wait_all_handles = staticmethod(_c.wait_all_handles)


ops = _Ops()


def import_op_lib():
    return ops


all_reduce_op_map = {t: _all_reduce for t in _HOST + _CUDA}
all_gather_op_map = {t: _all_gather for t in _HOST + _CUDA}
all_reduce_async_op_map = {t: _all_reduce_async for t in _HOST + _CUDA}  # the reference has the asynchronous forms on CUDA only
broadcast_async_op_map = {t: _broadcast_async for t in _HOST + _CUDA}
