"""Misc helpers (reference: srcs/python/kungfu/_utils.py:32-50 `_log_event`, `measure`, `show_duration`;
srcs/python/kungfu/utils/ema.py `EMA`; `kungfu.info`)."""
from __future__ import annotations

import os
import time
from typing import Callable, Tuple

from ..ops.state import ExponentialMovingAverage as EMA  # noqa: F401


def measure(f: Callable) -> Tuple[float, object]:
    t0 = time.time()
    r = f()
    return time.time() - t0, r


def show_duration(d: float) -> str:
    if d < 1e-3:
        return f"{d * 1e6:.2f}us"
    if d < 1:
        return f"{d * 1e3:.2f}ms"
    if d < 60:
        return f"{d:.2f}s"
    return f"{int(d) // 60}m{d % 60:.0f}s"


def show_rate(bytes_per_s: float) -> str:
    for unit, k in (("GiB/s", 1 << 30), ("MiB/s", 1 << 20), ("KiB/s", 1 << 10)):
        if bytes_per_s >= k:
            return f"{bytes_per_s / k:.2f}{unit}"
    return f"{bytes_per_s:.2f}B/s"


def map_maybe(f: Callable, lst):
    """Apply `f` to the non-None entries (reference: _utils.py:5-6)."""
    return [f(x) if x is not None else None for x in lst]


def one_based_range(n: int):
    return range(1, n + 1)


def _since_job_start() -> float:
    """Seconds since `kungfu-run` started the job (KUNGFU_JOB_START_TIMESTAMP; reference: _utils.py:32-36)."""
    t0 = float(os.environ.get("KUNGFU_JOB_START_TIMESTAMP", "0") or 0)
    return time.time() - t0 if t0 else 0.0


def _since_proc_start() -> float:
    t0 = float(os.environ.get("KUNGFU_PROC_START_TIMESTAMP", "0") or 0)
    return time.time() - t0 if t0 else 0.0


def _log_event(name: str) -> None:
    """`TS=... <name> :: <d> since proc started` using KUNGFU_PROC_START_TIMESTAMP (elastic join latency profiling)."""
    t0 = float(os.environ.get("KUNGFU_PROC_START_TIMESTAMP", "0") or 0)
    now = time.time()
    d = show_duration(now - t0) if t0 else "n/a"
    print(f"TS={now:.3f} {name} :: {d} since proc started", flush=True)


def info() -> dict:
    """Build / runtime information (reference: `python -m kungfu.info`)."""
    import torch

    from .. import __version__

    d = {"version": __version__, "torch": torch.__version__, "cuda_build": torch.version.cuda,
         "cuda_available": torch.cuda.is_available()}
    try:
        from .. import _kfrt  # noqa: F401

        d["native_runtime"] = True
    except ImportError:
        d["native_runtime"] = False
    try:
        from .. import _kfcuda

        d["cuda_kernels"] = True
        d["max_ranks"] = _kfcuda.MAX_RANKS
    except ImportError:
        d["cuda_kernels"] = False
    return d
