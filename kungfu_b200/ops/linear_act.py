"""Linear + bias (+ GELU) on the tcgen05 tensor cores with the activation fused into the GEMM epilogue
(csrc/kernels/gemm_tc5.cu: a hand-written persistent kernel — TMA-fed 4-stage ring, `tcgen05.mma` 128x256x16 issued by one
thread, two TMEM accumulators so that the bias+GELU epilogue of one tile overlaps the mainloop of the next).

`linear_act(x, weight, bias, act)` == `gelu(F.linear(x, weight, bias))` / `F.linear(...)` for bf16 CUDA tensors with
in_features % 64 == 0 and out_features % 8 == 0; anything else falls back to the PyTorch ops. Backward recomputes the
pre-activation (one library GEMM) instead of storing it.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

_ACT = {None: 0, "identity": 0, "gelu": 1}
_ws = {}


def _ext():
    from ..parallel.mesh import _load_ext

    return _load_ext()


def supported(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> bool:
    if not (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and bias is not None and bias.dtype == torch.bfloat16):
        return False
    if x.shape[-1] % 64 or weight.shape[0] % 8 or not weight.is_contiguous() or x.numel() == 0:
        return False
    try:
        return bool(_ext().gemm_sm100_available())
    except ImportError:
        return False


def _run(x2d: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, act: int) -> torch.Tensor:
    ext = _ext()
    tokens, in_f = x2d.shape
    out_f = weight.shape[0]
    y = torch.empty((tokens, out_f), dtype=torch.bfloat16, device=x2d.device)
    key = (x2d.device.index, tokens, in_f, out_f, act)
    if key not in _ws:
        n = ext.gemm_bias_act_workspace(tokens, in_f, out_f, act)
        _ws[key] = torch.empty(max(int(n), 16), dtype=torch.uint8, device=x2d.device)
    ws = _ws[key]
    ext.gemm_bias_act(x2d.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(), tokens, in_f, out_f, act, ws.data_ptr(), ws.numel(),
                      torch.cuda.current_stream().cuda_stream)
    return y


class _LinearAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):
        x2d = x.reshape(-1, x.shape[-1]).contiguous()
        y = _run(x2d, weight, bias, act)
        ctx.save_for_backward(x2d, weight, bias)
        ctx.act = act
        ctx.shape = x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2d, weight, bias = ctx.saved_tensors
        dz = dy.reshape(-1, dy.shape[-1])
        if ctx.act == 1:
            z = F.linear(x2d, weight, bias)  # recompute the pre-activation
            dz = torch.ops.aten.gelu_backward(dz, z)
        dx = (dz @ weight).view(ctx.shape)
        dw = dz.t() @ x2d
        db = dz.sum(0)
        return dx, dw, db, None


def linear_act(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], act: Optional[str] = None) -> torch.Tensor:
    code = _ACT[act]
    if supported(x, weight, bias):
        return _LinearAct.apply(x, weight, bias, code)
    y = F.linear(x, weight, bias)
    return F.gelu(y) if code == 1 else y


class FusedLinear(torch.nn.Linear):
    """nn.Linear whose forward runs the tcgen05 GEMM with the activation in the epilogue."""

    def __init__(self, in_features: int, out_features: int, act: Optional[str] = None) -> None:
        super().__init__(in_features, out_features, bias=True)
        self.act = act

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        return linear_act(x, self.weight, self.bias, self.act)
