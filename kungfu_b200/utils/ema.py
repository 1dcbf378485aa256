"""`kungfu.utils.ema` — the module path of the reference (srcs/python/kungfu/utils/ema.py: `EMA(alpha, scale_cap)` with
`update` / `get` / `reset`; a new sample is clamped to [value / scale_cap, value * scale_cap] before it is blended in)."""
from ..ops.state import ExponentialMovingAverage as EMA  # noqa: F401

__all__ = ["EMA"]
