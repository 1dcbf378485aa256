"""Fused transformer element-wise ops on our sm_100a kernels (csrc/kernels/transformer.cu):

    add_layernorm(x, residual, weight, bias, eps)  ==  F.layer_norm(x + residual, (H,), weight, bias, eps)
    bias_gelu(x, bias)                             ==  F.gelu(x + bias)

Forward is one pass each; backward produces the input gradient AND the parameter gradients (dgamma/dbeta, dbias) from the
same pass with a deterministic two-stage column reduction. bf16 / fp16 CUDA tensors with a last dimension that is a multiple
of 8 take the kernels; everything else (CPU, fp32, odd sizes) falls back to the PyTorch expression, so models built on these
ops also run in the CPU test-suite.

The reference ships no model code; these serve the BERT-large PairAveraging configuration of BASELINE.json.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from ..utils import dtypes


def _ext():
    from ..parallel.mesh import _load_ext

    return _load_ext()


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _fast(x: torch.Tensor, *others: Optional[torch.Tensor]) -> bool:
    if not (x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and x.shape[-1] % 8 == 0 and x.shape[-1] >= 8):
        return False
    return all(o is None or (o.is_cuda and o.dtype == x.dtype) for o in others)


class _AddLayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, eps):
        H = x.shape[-1]
        x2 = x.reshape(-1, H).contiguous()
        r2 = residual.reshape(-1, H).contiguous() if residual is not None else None
        rows = x2.shape[0]
        ext = _ext()
        y = torch.empty_like(x2)
        z = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        w, b = weight.contiguous(), bias.contiguous()
        ext.add_ln_fwd(x2.data_ptr(), r2.data_ptr() if r2 is not None else 0, w.data_ptr(), b.data_ptr(), z.data_ptr(), y.data_ptr(),
                       mean.data_ptr(), rstd.data_ptr(), dtypes.dtype_code(x.dtype), rows, H, float(eps), _st())
        ctx.save_for_backward(z, mean, rstd, w)
        ctx.has_res = residual is not None
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        z, mean, rstd, w = ctx.saved_tensors
        rows, H = z.shape
        ext = _ext()
        dy2 = dy.reshape(rows, H).contiguous()
        dz = torch.empty_like(z)
        dgamma = torch.empty_like(w)
        dbeta = torch.empty_like(w)
        partial = torch.empty(ext.add_ln_num_blocks(rows) * 2 * H, dtype=torch.float32, device=z.device)
        ext.add_ln_bwd(dy2.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), w.data_ptr(), dz.data_ptr(), dgamma.data_ptr(),
                       dbeta.data_ptr(), partial.data_ptr(), dtypes.dtype_code(z.dtype), rows, H, _st())
        dz = dz.view(ctx.shape)
        return dz, (dz if ctx.has_res else None), dgamma, dbeta, None


def add_layernorm(x: torch.Tensor, residual: Optional[torch.Tensor], weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm over the last dimension of `x + residual` (residual may be None)."""
    if _fast(x, residual, weight, bias) and x.shape[-1] <= 2048:
        return _AddLayerNorm.apply(x, residual, weight, bias, eps)
    z = x if residual is None else x + residual
    return F.layer_norm(z, (x.shape[-1],), weight, bias, eps)


class _BiasGelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias):
        I = x.shape[-1]
        x2 = x.reshape(-1, I).contiguous()
        b = bias.contiguous()
        y = torch.empty_like(x2)
        _ext().bias_gelu_fwd(x2.data_ptr(), b.data_ptr(), y.data_ptr(), dtypes.dtype_code(x.dtype), x2.shape[0], I, _st())
        ctx.save_for_backward(x2, b)
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, b = ctx.saved_tensors
        rows, I = x2.shape
        ext = _ext()
        dy2 = dy.reshape(rows, I).contiguous()
        dx = torch.empty_like(x2)
        db = torch.empty_like(b)
        partial = torch.empty(ext.bias_gelu_num_blocks(rows) * I, dtype=torch.float32, device=x2.device)
        ext.bias_gelu_bwd(dy2.data_ptr(), x2.data_ptr(), b.data_ptr(), dx.data_ptr(), db.data_ptr(), partial.data_ptr(), dtypes.dtype_code(x2.dtype),
                          rows, I, _st())
        return dx.view(ctx.shape), db


def bias_gelu(x: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """gelu(x + bias), exact (erf) form, bias broadcast over the last dimension."""
    if _fast(x, bias) and x.shape[-1] <= 16384:
        return _BiasGelu.apply(x, bias)
    return F.gelu(x + bias)
