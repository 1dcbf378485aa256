"""SynchronousSGDOptimizer (S-SGD).

Reference: TF `SynchronousSGDOptimizer` (srcs/python/kungfu/tensorflow/optimizers/sync_sgd.py:15-109: all-reduce-sum every
gradient, divide by np, apply) and the torch stub (srcs/python/kungfu/torch/optimizers/sync_sgd.py:6-32: D2H -> CPU
all-reduce -> H2D per tensor after backward, no overlap, no averaging).

Here, on GPUs:
  * fused path (torch.optim.SGD / Adam / AdamW): per bucket ONE sm_100a kernel does
    reduce-scatter(grads) -> optimizer on the owned shard -> all-gather(params) over NVLink (P2P or NVLS),
    launched from autograd hooks on the comm stream, overlapping the rest of backward;
  * generic path (any other optimizer): per-bucket in-place averaging all-reduce kernel, then the optimizer's own step.
On CPU tensors it behaves like the reference (named per-tensor all-reduce over the CPU runtime).
"""
from __future__ import annotations

from typing import Optional

import torch

from ...utils import dtypes
from .core import KungFuMixin, cpu_all_reduce_grads, fusable_kind, wrap_optimizer


class _SynchronousSGD(KungFuMixin):
    def _kf_init_ssgd(self, named_parameters, op, fused, overlap, bucket_mb, monitor, algo, use_gpu_engine=True, backend=None):
        self._kf_setup(named_parameters, use_gpu_engine)
        code, avg = dtypes.op_code(op if op is not None else "avg")
        if code != dtypes.SUM:
            raise ValueError("SynchronousSGDOptimizer reduces gradients with sum or avg")
        self._kf_average = avg
        self._kf_op = "sum"
        self._kf_monitor = monitor
        kind = fusable_kind(self) if fused in (True, "auto") else None
        if fused is True and kind is None:
            raise ValueError("kungfu_b200: fused=True but this optimizer has no fused kernel (SGD/Adam/AdamW only)")
        self._kf_kind = kind
        if self._kf_gpu:
            from ...parallel.engine import FlatEngine

            self._kf_engine = FlatEngine(self.param_groups, self._kf_named, mode="ssgd" if kind else "grads", opt_kind=kind,
                                         average=avg, overlap=overlap, bucket_mb=bucket_mb, monitor=monitor, algo=algo,
                                         name="ssgd", backend=backend)

    def sync_gradients(self):
        """Average (or sum) gradients across workers without stepping (generic path only)."""
        if self._kf_engine is not None:
            if self._kf_engine.mode != "grads":
                raise RuntimeError("sync_gradients() is not available on the fused path: step() does both")
            self._kf_engine.step_buckets()
        else:
            cpu_all_reduce_grads(self._kf_named, self._kf_op, self._kf_average)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        eng = self._kf_engine
        if eng is None:
            cpu_all_reduce_grads(self._kf_named, self._kf_op, self._kf_average)
            self._kf_inner_step()
        elif eng.mode == "ssgd":
            eng.step_buckets()
        else:
            eng.step_buckets()
            self._kf_inner_step()
        self._kf_step_count += 1
        return loss


def SynchronousSGDOptimizer(optimizer: torch.optim.Optimizer, named_parameters=None, op: Optional[str] = None,
                            fused="auto", overlap: bool = True, bucket_mb: Optional[float] = None, monitor: bool = False,
                            algo: Optional[str] = None, backend: Optional[str] = None):
    """Wrap `optimizer` so that `step()` applies the gradient averaged over all workers.

    `op`: None/'avg' (default, matches the TF optimizer and Horovod) or 'sum' (the reference torch binding).
    `overlap=True` (default) reduces — and on the fused path applies — every bucket from autograd hooks while backward is
    still running. Consequences: gradient accumulation needs `with opt.no_sync():` around all but the last backward;
    `clip_grad_norm_` / AMP `unscale_` between backward() and step() see gradients that have already been consumed — build
    the optimizer with `overlap=False` for those recipes (everything then happens inside step(), as in the reference).
    `backend='nccl'`: same engine, buckets and local optimizer kernels, but the collective is `ncclAllReduce` (baseline arm).
    """
    return wrap_optimizer(optimizer, _SynchronousSGD,
                          lambda o: o._kf_init_ssgd(named_parameters, op, fused, overlap, bucket_mb, monitor, algo, backend=backend))
