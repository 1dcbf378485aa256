"""MonitorGradientVarianceOptimizer: S-SGD + gradient variance across workers.

Reference (TF only): srcs/python/kungfu/tensorflow/optimizers/grad_variance.py:9-75 — an extra all-reduce of g^2 gives
E[g^2]; variance = E[g^2] - E[g]^2, and the sum of its norm is printed. Here sum_i Var_i = mean_r |g_r|^2 - |g_avg|^2,
which the fused kernels already produce as side reductions (no second all-reduce, K11 in SURVEY.md §2.3).
"""
from __future__ import annotations

import torch

from ...ops import collective as coll
from .core import wrap_optimizer
from .sync_sgd import _SynchronousSGD


class _MonitorVariance(_SynchronousSGD):
    def _kf_init_var(self, named_parameters, monitor_interval, fused, overlap, bucket_mb, algo):
        self._kf_init_ssgd(named_parameters, "avg", fused, overlap, bucket_mb, True, algo)
        self._kf_interval = int(monitor_interval)
        self.variance_value = 0.0
        self._kf_var_dev = None

    @torch.no_grad()
    def step(self, closure=None):
        n = self._kf_rt.size
        eng = self._kf_engine
        if eng is None:
            local_sq = sum(float((p.grad.detach().float() ** 2).sum()) for _, p in self._kf_named if p.grad is not None)
            loss = _SynchronousSGD.step(self, closure)
            big_sq = sum(float((p.grad.detach().float() ** 2).sum()) for _, p in self._kf_named if p.grad is not None)
            t = torch.tensor([local_sq], dtype=torch.float64)
            coll.all_reduce_(t, "sum")
            self.variance_value = float(t.item()) / n - big_sq
            return loss
        loss = _SynchronousSGD.step(self, closure)
        sums = eng.read_sumsq(reset=True)
        self._kf_var_dev = sums[0] / n - sums[1]
        return loss

    def variance(self) -> float:
        """Sum over all gradient elements of the across-worker variance."""
        if self._kf_var_dev is not None:
            return float(self._kf_var_dev.item())
        return float(self.variance_value)


def MonitorGradientVarianceOptimizer(optimizer, named_parameters=None, monitor_interval: int = 1, fused="auto", overlap: bool = True,
                                     bucket_mb=None, algo=None):
    return wrap_optimizer(optimizer, _MonitorVariance,
                          lambda o: o._kf_init_var(named_parameters, monitor_interval, fused, overlap, bucket_mb, algo))
