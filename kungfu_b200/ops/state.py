"""Small stateful ops: counter and exponential moving average.

Reference: `KungfuCounter`, `KungfuExponentialMovingAverage` (srcs/cpp/src/tensorflow/ops/cpu/state.cpp; EMA rule in
srcs/cpp/include/kungfu/utils/ema.hpp:19-28: first sample passes through, then v = alpha*v + (1-alpha)*x) and the python
helper with `scale_cap` clamping (srcs/python/kungfu/utils/ema.py:7-24).
"""
from __future__ import annotations

from typing import Optional


class Counter:
    def __init__(self, init: int = 0, incr: int = 1) -> None:
        self.value = int(init)
        self.incr = int(incr)

    def __call__(self) -> int:
        v = self.value
        self.value += self.incr
        return v


def counter(init: Optional[int] = None, incr: Optional[int] = None) -> Counter:
    return Counter(0 if init is None else init, 1 if incr is None else incr)


class ExponentialMovingAverage:
    def __init__(self, alpha: float = 0.9, scale_cap: Optional[float] = None) -> None:
        self.alpha = float(alpha)
        self.scale_cap = scale_cap
        self.value: Optional[float] = None

    def update(self, x: float) -> float:
        x = float(x)
        if self.value is None:
            self.value = x
        else:
            if self.scale_cap is not None and self.value != 0:
                lo, hi = sorted((self.value / self.scale_cap, self.value * self.scale_cap))
                x = min(max(x, lo), hi)
            self.value = self.alpha * self.value + (1.0 - self.alpha) * x
        return self.value

    __call__ = update

    def get(self) -> float:
        return 0.0 if self.value is None else self.value

    def reset(self) -> None:
        """Forget the history: the next sample is taken as it is (reference: utils/ema.py `reset`)."""
        self.value = None


def exponential_moving_average(alpha: float = 0.9) -> ExponentialMovingAverage:
    return ExponentialMovingAverage(alpha)
