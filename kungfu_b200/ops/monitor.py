"""Monitoring ops: egress rates and the gradient-noise-scale estimator.

Reference: srcs/python/kungfu/tensorflow/ops/monitor.py:6-27 (`global_noise_scale`, `egress_rates`), kernel
`KungfuNoiseScale` (srcs/cpp/src/tensorflow/ops/cpu/collective.cpp:212-258).
"""
from __future__ import annotations

import torch

from ..runtime import get_runtime
from .state import ExponentialMovingAverage


def egress_rates() -> torch.Tensor:
    """float32[np]: bytes/s sent to every peer since the previous call (host transport; on the GPU plane traffic is
    NVLink loads/stores issued by kernels and is reported per op by the benchmarks instead)."""
    rt = get_runtime()
    f = getattr(rt.control, "egress_rates", None)
    return torch.tensor(f() if f else [0.0] * rt.size, dtype=torch.float32)


class NoiseScale:
    """G_biased = (B_big |g_big|^2 - B_small |g_small|^2) / (B_big - B_small)
       S_biased = (|g_small|^2 - |g_big|^2) / (1/B_small - 1/B_big);  noise scale = EMA(S) / EMA(G)."""

    def __init__(self, alpha: float = 0.6) -> None:
        self.g_ema = ExponentialMovingAverage(alpha)
        self.s_ema = ExponentialMovingAverage(alpha)
        self.last_g = 0.0
        self.last_s = 0.0

    def update(self, batch_small: float, batch_big: float, g_small_sq: float, g_big_sq: float) -> float:
        if batch_big == batch_small:
            return 0.0
        self.last_g = (batch_big * g_big_sq - batch_small * g_small_sq) / (batch_big - batch_small)
        self.last_s = (g_small_sq - g_big_sq) / (1.0 / batch_small - 1.0 / batch_big)
        g = self.g_ema.update(self.last_g)
        s = self.s_ema.update(self.last_s)
        return s / g if g != 0 else 0.0


def global_noise_scale(batch_small, batch_big, g_small_sq, g_big_sq, alpha: float = 0.6, _state={}):
    ns = _state.setdefault(alpha, NoiseScale(alpha))
    return ns.update(float(batch_small), float(batch_big), float(g_small_sq), float(g_big_sq))
