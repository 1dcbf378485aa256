"""SynchronousAveragingOptimizer (SMA / EA-SGD).

Reference (TF only): srcs/python/kungfu/tensorflow/optimizers/sma_sgd.py:9-74 — every step all-reduce-sum the *model*,
`avg = sum / np`, `v <- (1 - alpha) v + alpha avg` (alpha = 0.1), THEN apply the local gradients.

GPU path (overlap=True, default): the model average does not depend on this step's gradients, so `zero_grad()` starts an
out-of-place split-phase all-reduce (params -> scratch, 1/np folded in) on the comm stream and the whole transfer hides
under forward+backward; after backward ONE local kernel does blend + SGD/Adam step over the arena. overlap=False (or no
`zero_grad()` before backward): `sma_fused_kernel` does reduce-scatter -> 1/np -> all-gather -> blend -> step in a single
kernel after backward (csrc/kernels/fused_step.cu).
"""
from __future__ import annotations

from typing import Optional

import torch

from ...ops import collective as coll
from .core import KungFuMixin, fusable_kind, wrap_optimizer


class _SynchronousAveraging(KungFuMixin):
    def _kf_init_sma(self, named_parameters, alpha, fused, bucket_mb, algo, overlap=True, backend=None):
        self._kf_setup(named_parameters, True)
        self._kf_alpha = float(alpha)
        kind = fusable_kind(self) if fused in (True, "auto") else None
        if fused is True and kind is None:
            raise ValueError("kungfu_b200: fused=True but this optimizer has no fused kernel (SGD/Adam/AdamW only)")
        self._kf_kind = kind
        if self._kf_gpu and kind is not None:
            from ...parallel.engine import FlatEngine

            self._kf_engine = FlatEngine(self.param_groups, self._kf_named, mode="sma", opt_kind=kind, average=False,
                                         overlap=overlap, bucket_mb=bucket_mb, alpha=self._kf_alpha, algo=algo, name="sma",
                                         backend=backend)

    def _kf_begin_step(self) -> None:
        if self._kf_engine is not None:
            self._kf_engine.begin_step()

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
                                  bucket_mb: Optional[float] = None, algo: Optional[str] = None, overlap: bool = True,
                                  backend: Optional[str] = None):
    """Wrap `optimizer` with synchronous model averaging (reference signature: `(optimizer, name, alpha=0.1)`)."""
    return wrap_optimizer(optimizer, _SynchronousAveraging,
                          lambda o: o._kf_init_sma(named_parameters, alpha, fused, bucket_mb, algo, overlap, backend))
