"""SynchronousAveragingOptimizer (SMA / EA-SGD).

Reference (TF only): srcs/python/kungfu/tensorflow/optimizers/sma_sgd.py:9-74 — every step all-reduce-sum the *model*,
`avg = sum / np`, `v <- (1 - alpha) v + alpha avg` (alpha = 0.1), THEN apply the local gradients.

GPU fused path: one `sma_fused_kernel` per bucket does reduce-scatter(params) -> 1/np -> all-gather into a scratch region
-> blend -> local SGD/Adam step in a single pass over the arena (csrc/kernels/fused_step.cu).
"""
from __future__ import annotations

from typing import Optional

import torch

from ...ops import collective as coll
from .core import KungFuMixin, fusable_kind, wrap_optimizer


class _SynchronousAveraging(KungFuMixin):
    def _kf_init_sma(self, named_parameters, alpha, fused, bucket_mb, algo):
        self._kf_setup(named_parameters, True)
        self._kf_alpha = float(alpha)
        kind = fusable_kind(self) if fused in (True, "auto") else None
        if fused is True and kind is None:
            raise ValueError("kungfu_b200: fused=True but this optimizer has no fused kernel (SGD/Adam/AdamW only)")
        self._kf_kind = kind
        if self._kf_gpu and kind is not None:
            from ...parallel.engine import FlatEngine

            self._kf_engine = FlatEngine(self.param_groups, self._kf_named, mode="sma", opt_kind=kind, average=False,
                                         overlap=False, bucket_mb=bucket_mb, alpha=self._kf_alpha, algo=algo, name="sma")

    def average_model(self):
        """v <- (1 - alpha) v + alpha mean_over_workers(v) (generic path)."""
        rt = self._kf_rt
        if rt.size == 1:
            return
        a = self._kf_alpha
        handles, avgs = [], []
        for name, p in self._kf_named:
            avg = p.data.detach().clone()
            if avg.is_cuda:
                handles.append(coll.all_reduce_async_(avg, name, "avg"))
            else:
                coll.all_reduce_(avg, "avg", name)
            avgs.append(avg)
        coll.wait_all_handles(handles)
        for (name, p), avg in zip(self._kf_named, avgs):
            p.data.mul_(1.0 - a).add_(avg, alpha=a)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        eng = self._kf_engine
        if eng is not None:
            eng.step_buckets()
        else:
            self.average_model()
            self._kf_inner_step()
        self._kf_step_count += 1
        return loss


def SynchronousAveragingOptimizer(optimizer: torch.optim.Optimizer, named_parameters=None, alpha: float = 0.1, fused="auto",
                                  bucket_mb: Optional[float] = None, algo: Optional[str] = None):
    """Wrap `optimizer` with synchronous model averaging (reference signature: `(optimizer, name, alpha=0.1)`)."""
    return wrap_optimizer(optimizer, _SynchronousAveraging, lambda o: o._kf_init_sma(named_parameters, alpha, fused, bucket_mb, algo))
