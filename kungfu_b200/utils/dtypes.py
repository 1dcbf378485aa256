"""dtype / reduce-op codes shared by the Python layer, the CPU runtime and the CUDA kernels.

Mirrors the reference's `KungFu_Datatype` / `KungFu_Op` encoders (srcs/cpp/include/kungfu/dtype.h:21-39,
op.h:8-19) but with bf16 as a first-class type (the reference maps bf16 onto fp16: tensorflow/ops.h:21-22).
"""
from __future__ import annotations

import torch

F32, BF16, F16, I32, I64, U8, F64, I8, U32, U64, I16, U16 = range(12)

_TORCH2CODE = {
    torch.float32: F32,
    torch.bfloat16: BF16,
    torch.float16: F16,
    torch.int32: I32,
    torch.int64: I64,
    torch.uint8: U8,
    torch.float64: F64,
    torch.int8: I8,
    torch.int16: I16,
    torch.bool: U8,
}
for _name, _code in (("uint16", U16), ("uint32", U32), ("uint64", U64)):
    if hasattr(torch, _name):
        _TORCH2CODE[getattr(torch, _name)] = _code

_SIZE = {F32: 4, BF16: 2, F16: 2, I32: 4, I64: 8, U8: 1, F64: 8, I8: 1, U32: 4, U64: 8, I16: 2, U16: 2}

SUM, MIN, MAX, PROD = range(4)
_OPS = {"sum": SUM, "min": MIN, "max": MAX, "prod": PROD}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _TORCH2CODE[dt]
    except KeyError:
        raise TypeError(f"kungfu_b200: unsupported dtype {dt}") from None


def dtype_size(code: int) -> int:
    return _SIZE[code]


def op_code(op) -> tuple[int, bool]:
    """Returns (reduce op code, average?). `op` may be None/'sum'/'min'/'max'/'prod'/'avg'/'mean'."""
    if op is None:
        return SUM, False
    if isinstance(op, int):
        return op, False
    name = str(op).lower()
    if name in ("avg", "mean", "average"):
        return SUM, True
    if name not in _OPS:
        raise ValueError(f"kungfu_b200: unknown reduce op {op!r}")
    return _OPS[name], False
