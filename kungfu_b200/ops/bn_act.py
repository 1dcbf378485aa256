"""Fused BatchNorm (+ residual add) + ReLU for channels_last bf16/fp16 activations (csrc/kernels/bn_act.cu).

`bn_act(x, weight, bias, running_mean, running_var, residual=None, relu=True, training=True, momentum=0.1, eps=1e-5)`
is a drop-in for `relu(batch_norm(x) + residual)`; backward needs only (dy, x, 1-bit mask). Falls back to the PyTorch
ops for layouts / dtypes / channel counts the kernels do not cover (CPU tensors, fp32 activations, C % 8 != 0).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from ..utils import dtypes

_ext = None


def _get_ext():
    global _ext
    if _ext is None:
        from ..parallel.mesh import _load_ext

        _ext = _load_ext()
    return _ext


def supported(x: torch.Tensor, weight: Optional[torch.Tensor]) -> bool:
    if not x.is_cuda or x.dim() != 4 or x.dtype not in (torch.bfloat16, torch.float16):
        return False
    C = x.shape[1]
    cvn = C // 8
    if C % 8 != 0 or not ((cvn <= 256 and 256 % cvn == 0) or (cvn > 256 and cvn % 256 == 0)) or C > 8192:
        return False
    if weight is not None and weight.dtype not in (torch.bfloat16, torch.float16, torch.float32):
        return False
    return x.is_contiguous(memory_format=torch.channels_last)


class _BNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, training, momentum, eps, relu):
        ext = _get_ext()
        N, C, H, W = x.shape
        rows = N * H * W
        dev = x.device
        st = torch.cuda.current_stream().cuda_stream
        y = torch.empty_like(x, memory_format=torch.channels_last)
        mask = torch.empty(rows * C // 8, dtype=torch.uint8, device=dev) if relu else None
        dt = dtypes.dtype_code(x.dtype)
        pdt = dtypes.dtype_code(weight.dtype)
        res = residual if residual is not None else None
        p = lambda t: t.data_ptr() if t is not None else 0
        if training:
            nb = ext.bn_num_blocks(rows, C)
            partial = torch.empty(nb * 2 * C, dtype=torch.float32, device=dev)
            stats = torch.empty(4, C, dtype=torch.float32, device=dev)  # mean, invstd, scale, shift
            ext.bn_fwd(p(x), p(res), p(y), p(mask), dt, rows, C, p(weight), p(bias), p(running_mean), p(running_var), pdt, float(eps),
                       float(momentum), int(relu), p(partial), stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(),
                       stats[3].data_ptr(), st)
            ctx.save_for_backward(x, mask, weight, stats)
        else:
            scale = weight.float() * torch.rsqrt(running_var.float() + eps)
            shift = bias.float() - running_mean.float() * scale
            ext.bn_apply(p(x), p(res), p(y), p(mask), dt, rows, C, scale.data_ptr(), shift.data_ptr(), int(relu), st)
            stats = torch.stack([running_mean.float(), torch.rsqrt(running_var.float() + eps), scale, shift])
            ctx.save_for_backward(x, mask, weight, stats)
        ctx.relu = relu
        ctx.training = training
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        ext = _get_ext()
        x, mask, weight, stats = ctx.saved_tensors
        N, C, H, W = x.shape
        rows = N * H * W
        dev = x.device
        st = torch.cuda.current_stream().cuda_stream
        dy = dy.contiguous(memory_format=torch.channels_last)
        p = lambda t: t.data_ptr() if t is not None else 0
        if not ctx.training:
            # eval mode: y = relu(x * scale + shift + res): plain chain rule, no batch statistics involved
            dz = dy
            if ctx.relu:
                bits = ((mask[:, None] >> torch.arange(8, device=dev, dtype=torch.uint8)) & 1).view(N, H, W, C).permute(0, 3, 1, 2)
                dz = dy * bits.to(dy.dtype)
            dx = dz * stats[2].to(dy.dtype).view(1, C, 1, 1)
            xhat = (x.float() - stats[0].view(1, C, 1, 1)) * stats[1].view(1, C, 1, 1)
            dw = (dz.float() * xhat).sum((0, 2, 3)).to(weight.dtype)
            db = dz.float().sum((0, 2, 3)).to(weight.dtype)
            return dx, (dz if ctx.has_res else None), dw, db, None, None, None, None, None, None
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        dres = torch.empty_like(x, memory_format=torch.channels_last) if ctx.has_res else None
        dw = torch.empty_like(weight)
        db = torch.empty_like(weight)
        nb = ext.bn_num_blocks(rows, C)
        partial = torch.empty(nb * 2 * C, dtype=torch.float32, device=dev)
        coef = torch.empty(3 * C, dtype=torch.float32, device=dev)
        ext.bn_bwd(p(dy), p(x), p(mask), p(dx), p(dres), dtypes.dtype_code(x.dtype), rows, C, p(weight), p(dw), p(db),
                   dtypes.dtype_code(weight.dtype), stats[0].data_ptr(), stats[1].data_ptr(), int(ctx.relu), p(partial), p(coef), st)
        return dx, dres, dw, db, None, None, None, None, None, None


def bn_act(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor],
           residual: Optional[torch.Tensor] = None, relu: bool = True, training: bool = True, momentum: float = 0.1, eps: float = 1e-5) -> torch.Tensor:
    if supported(x, weight) and (residual is None or (residual.shape == x.shape and residual.dtype == x.dtype)) and \
            (training or running_mean is not None):
        if residual is not None and not residual.is_contiguous(memory_format=torch.channels_last):
            residual = residual.contiguous(memory_format=torch.channels_last)
        return _BNAct.apply(x, residual, weight, bias, running_mean, running_var, training, momentum, eps, relu)
    y = F.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y
