"""`kungfu.torch.ops.collective` — the module path of the reference (srcs/python/kungfu/torch/ops/collective.py:8-55); the functions
live in the package `kungfu_b200.torch.ops`."""
from . import (Handle, all_gather, all_reduce_fn, barrier, broadcast_parameters, consensus, inplace_all_reduce_async_op,  # noqa: F401
               inplace_all_reduce_op, inplace_broadcast_async_op, inplace_broadcast_op, wait_all_handles, wait_handle)
