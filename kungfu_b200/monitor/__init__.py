"""Monitoring front end: network counters of the host plane, the `/metrics` endpoint of every peer, noise-scale estimators.

Reference: srcs/go/monitor/{counters,monitor,server}.go (per-peer egress / ingress byte counters and rates, Prometheus-style text
at `http://<peer>:<port+10000>/metrics`, enabled by KUNGFU_CONFIG_ENABLE_MONITORING), `egress_rates` op, and the gradient
noise-scale / variance monitors of the optimizers package.
"""
from __future__ import annotations

import os
import re
from typing import Dict, Optional

from ..ops.monitor import NoiseScale, egress_rates, global_noise_scale  # noqa: F401

_PLAIN = re.compile(r'^(?P<metric>[A-Za-z_][A-Za-z0-9_]*(?:\{[^}]*\})?)\s+(?P<value>[0-9.eE+-]+|nan|inf|-inf)\s*$')
_LINE = re.compile(r'^(?P<metric>[a-z_]+)\{peer="(?P<peer>[^"]+)"\}\s+(?P<value>[0-9.eE+-]+)\s*$')


def metrics_url(peer: Optional[str] = None) -> str:
    """`http://ip:port+10000/metrics` of `peer` ("ip:port"; default: this worker, from KUNGFU_SELF_SPEC)."""
    spec = peer or os.environ.get("KUNGFU_SELF_SPEC", "127.0.0.1:10000")
    ip, port = spec.rsplit(":", 1)
    return f"http://{ip}:{int(port) + 10000}/metrics"


def parse_metrics(text: str) -> Dict[str, Dict[str, float]]:
    """{'egress_total_bytes': {'ip:port': bytes, ...}, 'ingress_total_bytes': {...}}"""
    out: Dict[str, Dict[str, float]] = {}
    for line in text.splitlines():
        m = _LINE.match(line.strip())
        if m:
            out.setdefault(m.group("metric"), {})[m.group("peer")] = float(m.group("value"))
            continue
        m = _PLAIN.match(line.strip())
        if m:  # application gauges / counters: {'kungfu_cluster_size': {'': 4.0}, 'gpu_collective_bytes_total{op="all_reduce"}': {'': ...}}
            out.setdefault(m.group("metric"), {})[""] = float(m.group("value"))
    return out


def set_gauge(name: str, value: float) -> None:
    """Publish a gauge on this worker's `/metrics` (no-op without the native runtime)."""
    try:
        from .. import _kfrt  # type: ignore[attr-defined]

        _kfrt.metric_set(name, float(value))
    except ImportError:
        pass


def add_counter(name: str, delta: float) -> None:
    try:
        from .. import _kfrt  # type: ignore[attr-defined]

        _kfrt.metric_add(name, float(delta))
    except ImportError:
        pass


def fetch_metrics(peer: Optional[str] = None) -> Dict[str, Dict[str, float]]:
    """GET and parse a peer's `/metrics` (the job must run with KUNGFU_CONFIG_ENABLE_MONITORING=true)."""
    from .. import _kfrt  # type: ignore[attr-defined]

    status, body = _kfrt.http_request("GET", metrics_url(peer), "")
    if status != 200:
        raise RuntimeError(f"metrics endpoint {metrics_url(peer)} answered {status}")
    return parse_metrics(body.decode() if isinstance(body, (bytes, bytearray)) else body)
