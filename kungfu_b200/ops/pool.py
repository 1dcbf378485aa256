"""3x3 / stride 2 / pad 1 max pooling on channels_last bf16/fp16 tensors (csrc/kernels/pool.cu), forward + backward.
Falls back to `F.max_pool2d` for anything else."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ..utils import dtypes


def _ext():
    from ..parallel.mesh import _load_ext

    return _load_ext()


class _MaxPool3x3s2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        N, C, H, W = x.shape
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, C, OH, OW), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        idx = torch.empty(N * OH * OW * C, dtype=torch.uint8, device=x.device)
        _ext().maxpool3x3s2_fwd(x.data_ptr(), y.data_ptr(), idx.data_ptr(), dtypes.dtype_code(x.dtype), N, H, W, C,
                                torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(idx)
        ctx.shape = (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        N, C, H, W = ctx.shape
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, C, H, W), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        _ext().maxpool3x3s2_bwd(dy.data_ptr(), idx.data_ptr(), dx.data_ptr(), dtypes.dtype_code(dy.dtype), N, H, W, C,
                                torch.cuda.current_stream().cuda_stream)
        return dx


def max_pool3x3s2(x: torch.Tensor) -> torch.Tensor:
    if (x.is_cuda and x.dim() == 4 and x.dtype in (torch.bfloat16, torch.float16) and x.shape[1] % 8 == 0
            and x.is_contiguous(memory_format=torch.channels_last)):
        return _MaxPool3x3s2.apply(x)
    return F.max_pool2d(x, 3, 2, 1)


class MaxPool3x3s2(torch.nn.Module):
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return max_pool3x3s2(x)
