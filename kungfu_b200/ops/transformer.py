"""Fused transformer element-wise ops on our sm_100a kernels (csrc/kernels/transformer.cu):

    add_layernorm(x, residual, weight, bias, eps, lin_bias=None)  ==  F.layer_norm(x + lin_bias + residual, (H,), weight, bias, eps)
    bias_gelu(x, bias)                             ==  F.gelu(x + bias)
    qkv_projection(x, weight, bias, heads)         ==  F.linear(x, weight, bias) split into (q, k, v) views of shape [B, heads, S, d]

`lin_bias` is the bias of the GEMM that produced `x` (call the GEMM without it): the add is free in the forward pass and its gradient
is one more column sum in the backward pass the LayerNorm runs anyway. `qkv_projection` exists for its backward: attention hands back
three separate strided gradients, autograd would stack them (`cat`), copy the result into the projection's layout and run a column-sum
kernel for the bias; here ONE kernel writes the contiguous [tokens, 3H] gradient and the bias gradient.

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
    def forward(ctx, x, residual, weight, bias, eps, lin_bias):
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
        lb = lin_bias.contiguous() if lin_bias is not None else None
        ext.add_ln_fwd(x2.data_ptr(), r2.data_ptr() if r2 is not None else 0, lb.data_ptr() if lb is not None else 0, w.data_ptr(), b.data_ptr(),
                       z.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dtypes.dtype_code(x.dtype), rows, H, float(eps), _st())
        ctx.save_for_backward(z, mean, rstd, w)
        ctx.has_res = residual is not None
        ctx.has_lb = lin_bias is not None
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
        dlb = torch.empty_like(w) if ctx.has_lb else None
        partial = torch.empty(ext.add_ln_num_blocks(rows) * (3 if ctx.has_lb else 2) * H, dtype=torch.float32, device=z.device)
        ext.add_ln_bwd(dy2.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), w.data_ptr(), dz.data_ptr(), dgamma.data_ptr(),
                       dbeta.data_ptr(), dlb.data_ptr() if dlb is not None else 0, partial.data_ptr(), dtypes.dtype_code(z.dtype), rows, H, _st())
        dz = dz.view(ctx.shape)
        return dz, (dz if ctx.has_res else None), dgamma, dbeta, None, dlb


def add_layernorm(x: torch.Tensor, residual: Optional[torch.Tensor], weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5,
                  lin_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm over the last dimension of `x + lin_bias + residual` (residual / lin_bias may be None)."""
    if _fast(x, residual, weight, bias, lin_bias) and x.shape[-1] <= 2048:
        return _AddLayerNorm.apply(x, residual, weight, bias, eps, lin_bias)
    z = x if lin_bias is None else x + lin_bias
    z = z if residual is None else z + residual
    return F.layer_norm(z, (x.shape[-1],), weight, bias, eps)


class _QKVProjection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, heads):
        B, S, H = x.shape
        qkv = F.linear(x, weight, bias).view(B, S, 3, heads, H // heads)
        ctx.save_for_backward(x, weight)
        ctx.heads = heads
        ctx.has_bias = bias is not None
        return qkv[:, :, 0].transpose(1, 2), qkv[:, :, 1].transpose(1, 2), qkv[:, :, 2].transpose(1, 2)

    @staticmethod
    def backward(ctx, dq, dk, dv):
        x, weight = ctx.saved_tensors
        B, S, H = x.shape
        heads, d = ctx.heads, H // ctx.heads
        ext = _ext()
        gs = []
        for g in (dq, dk, dv):
            if g is None:
                g = torch.zeros(B, heads, S, d, dtype=x.dtype, device=x.device)
            if g.stride(-1) != 1 or any(st % 8 for st in g.stride()[:3]) or g.data_ptr() % 16:
                g = g.contiguous()
            gs.append(g)
        dqkv = torch.empty(B * S, 3 * H, dtype=x.dtype, device=x.device)
        dbias = torch.empty(3 * H, dtype=x.dtype, device=x.device)
        partial = torch.empty(ext.qkv_pack_num_blocks(B * S) * 3 * H, dtype=torch.float32, device=x.device)
        strides = [st for g in gs for st in (g.stride(0), g.stride(1), g.stride(2))]
        ext.qkv_pack_bwd(gs[0].data_ptr(), gs[1].data_ptr(), gs[2].data_ptr(), strides, dqkv.data_ptr(), dbias.data_ptr(), partial.data_ptr(),
                         dtypes.dtype_code(x.dtype), B, S, heads, d, _st())
        x2 = x.reshape(B * S, H)
        dx = (dqkv @ weight).view(B, S, H) if ctx.needs_input_grad[0] else None
        dw = dqkv.t() @ x2 if ctx.needs_input_grad[1] else None
        return dx, dw, (dbias if ctx.has_bias else None), None


def qkv_projection(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], heads: int):
    """Fused QKV projection of a [B, S, H] activation: returns (q, k, v), each a [B, heads, S, H / heads] view of one GEMM output."""
    B, S, H = x.shape
    d = H // heads
    if _fast(x, weight, bias) and d % 8 == 0 and 3 * H <= 8192 and weight.shape == (3 * H, H):
        return _QKVProjection.apply(x, weight, bias, heads)
    qkv = F.linear(x, weight, bias).view(B, S, 3, heads, d)
    return qkv[:, :, 0].transpose(1, 2), qkv[:, :, 1].transpose(1, 2), qkv[:, :, 2].transpose(1, 2)


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
