"""Scope tracing for Python-side code: NVTX ranges (visible in Nsight) + a per-scope summary + optional Chrome-trace JSON.

Reference: `TRACE_SCOPE` / stdtracer in C++ (srcs/cpp/include/kungfu/utils/trace.hpp:2-16; the native runtime has the
equivalent `KF_TRACE_SCOPE`), `_log_event`, `ResizeProfiler`, `LogPerfHook`. Enabled by
KUNGFU_ENABLE_TRACE=1; `KUNGFU_TRACE_FILE=path.json` additionally writes chrome://tracing events at exit (`path.<ip_port>.json` per worker under a launcher).
"""
from __future__ import annotations

import atexit
import json
import os
import time
from collections import defaultdict
from contextlib import contextmanager

ENABLED = os.environ.get("KUNGFU_ENABLE_TRACE", "0").lower() in ("1", "true", "yes", "on")
_stats = defaultdict(lambda: [0, 0.0])
_events = []


@contextmanager
def trace_scope(name: str, device_sync: bool = False):
    if not ENABLED:
        yield
        return
    nvtx = None
    try:
        import torch

        if torch.cuda.is_available():
            nvtx = torch.cuda.nvtx
            nvtx.range_push(name)
            if device_sync:
                torch.cuda.synchronize()
    except Exception:
        nvtx = None
    t0 = time.perf_counter()
    try:
        yield
    finally:
        if device_sync and nvtx is not None:
            import torch

            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if nvtx is not None:
            nvtx.range_pop()
        s = _stats[name]
        s[0] += 1
        s[1] += dt
        _events.append({"name": name, "ph": "X", "ts": t0 * 1e6, "dur": dt * 1e6, "pid": os.getpid(), "tid": 0})


def report() -> str:
    lines = ["[kungfu trace] %-40s %8s %14s %12s" % ("scope", "calls", "total(ms)", "mean(us)")]
    for k, (n, t) in sorted(_stats.items()):
        lines.append("[kungfu trace] %-40s %8d %14.3f %12.3f" % (k, n, t * 1e3, t * 1e6 / max(n, 1)))
    try:
        from .. import _kfrt

        native = _kfrt.trace_report()
        if native.count("\n") > 1:
            lines.append(native.rstrip())
    except Exception:
        pass
    return "\n".join(lines)


@atexit.register
def _dump() -> None:
    if not ENABLED or not _stats:
        return
    print(report(), flush=True)
    path = os.environ.get("KUNGFU_TRACE_FILE")
    if path:
        # one file per worker: under a launcher every process inherits the same KUNGFU_TRACE_FILE
        spec = os.environ.get("KUNGFU_SELF_SPEC") or os.environ.get("RANK")
        if spec:
            root, ext = os.path.splitext(path)
            path = f"{root}.{spec.replace(':', '_')}{ext}"
        with open(path, "w") as f:
            json.dump({"traceEvents": _events}, f)
