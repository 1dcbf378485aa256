"""Huawei ModelArts environment -> peer list.

Reference: srcs/go/platforms/modelarts/modelarts.go:14-115 — `DLS_TASK_INDEX`, `DLS_TASK_NUMBER` and
`BATCH_CUSTOM<i>_HOSTS` ("host:port") describe the containers of a job; every container runs `slots` workers.
Not needed on a single B200 box; provided for parity.
"""
from __future__ import annotations

import os
import socket
from typing import Dict, List, Optional, Tuple


def parse_env(env: Optional[Dict[str, str]] = None, slots: int = 8, port_begin: int = 10000) -> Tuple[int, List[str], List[str]]:
    """Returns (self container index, runner list ["ip:38080", ...], worker list ["ip:port", ...])."""
    env = dict(os.environ if env is None else env)
    n = int(env["DLS_TASK_NUMBER"])
    idx = int(env["DLS_TASK_INDEX"])
    hosts: List[str] = []
    for i in range(n):
        hp = env[f"BATCH_CUSTOM{i}_HOSTS"]
        host = hp.split(":")[0]
        try:
            host = socket.gethostbyname(host)
        except OSError:
            pass
        hosts.append(host)
    runners = [f"{h}:38080" for h in hosts]
    workers = [f"{h}:{port_begin + s}" for h in hosts for s in range(slots)]
    return idx, runners, workers


def kungfu_run_args(env: Optional[Dict[str, str]] = None, slots: int = 8) -> List[str]:
    """`kungfu-run` flags for this container: -np, -H and -self."""
    idx, runners, workers = parse_env(env, slots)
    hosts = ",".join(f"{r.split(':')[0]}:{slots}" for r in runners)
    return ["-np", str(len(workers)), "-H", hosts, "-self", runners[idx].split(":")[0]]
