"""`kungfu.cmd` parity: launcher entry point, failure-detector signals, in-process multi-worker launch.

Reference: srcs/python/kungfu/cmd/__init__.py:7-47 — `run()` (console script `kungfu-run`), `monitor_batch_begin /
monitor_batch_end / monitor_train_end / monitor_epoch_end` (POST {"key": "<kind>:<rank>"} to the first peer's IP :7756,
srcs/go/libkungfu-comm/send.go:16-57) and `launch_multiprocess(f, np)` (mp.Pool + SingleMachineEnv).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
from typing import Callable

MONITOR_PORT = int(os.environ.get("KUNGFU_MONITOR_PORT", "7756"))


def _bin(name: str) -> str:
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(here, "bin", name)
    if not os.path.exists(path):
        from .. import _build

        _build.build_runtime()
    return path


def run() -> None:
    """Console-script entry: exec the native `kungfu-run` launcher with the current argv."""
    exe = _bin("kungfu-run")
    os.execv(exe, [exe] + sys.argv[1:])


def kungfu_run_path() -> str:
    """Path of the native `kungfu-run` binary built in-tree."""
    return _bin("kungfu-run")


def config_server_path() -> str:
    """Path of the native `kungfu-config-server` binary."""
    return _bin("kungfu-config-server")


def _monitor_addr() -> str:
    peers = os.environ.get("KUNGFU_INIT_PEERS", "")
    host = peers.split(",")[0].rsplit(":", 1)[0] if peers else "127.0.0.1"
    # under `kungfu-run -w -shrink-on-failure -auto-recover d` every runner watches its own workers: report to it
    return os.environ.get("KUNGFU_MONITOR_ADDR", os.environ.get("KUNGFU_HANG_WATCH", f"http://{host}:{MONITOR_PORT}/"))


def _send_signal(kind: str) -> bool:
    from ..runtime import get_runtime

    try:
        from .. import _kfrt  # type: ignore[attr-defined]
    except ImportError:
        return False
    rank = get_runtime().rank
    status, _ = _kfrt.http_request("POST", _monitor_addr(), '{"key":"%s:%d"}' % (kind, rank))
    return status == 200


def monitor_batch_begin() -> bool:
    """Heartbeat: a batch started (the detector flags the job if no `end` follows within the threshold)."""
    return _send_signal("begin")


def monitor_batch_end() -> bool:
    """Heartbeat: the batch finished (clears the stall timer of this rank)."""
    return _send_signal("end")


def monitor_train_end() -> bool:
    """This rank finished training; when every rank has said so the launcher ends the job successfully."""
    return _send_signal("trainend")


def monitor_epoch_end() -> bool:
    """One more epoch is done (and checkpointed): after a restart `--n-epochs` is reduced by the minimum over ranks."""
    return _send_signal("epoch")


def _mp_entry(args) -> None:
    f, rank, size, base_port = args
    os.environ["KUNGFU_SELF_SPEC"] = f"127.0.0.1:{base_port + rank}"
    os.environ["KUNGFU_INIT_PEERS"] = ",".join(f"127.0.0.1:{base_port + i}" for i in range(size))
    os.environ.pop("KUNGFU_SINGLE_MACHINE_MULTIPROCESS", None)
    from ..runtime import get_runtime, reset_runtime

    reset_runtime()
    get_runtime()
    try:
        f(rank)
    finally:
        reset_runtime()


def launch_multiprocess(f: Callable[[int], None], np: int, base_port: int = 0) -> None:
    """Run `f(rank)` in `np` worker processes that form a KungFu cluster on this machine (no launcher).
    Reference: cmd/__init__.py:32-47."""
    if base_port == 0:
        base_port = 24000 + (os.getpid() * 13) % 20000
    ctx = mp.get_context("spawn")
    with ctx.Pool(np) as pool:
        pool.map(_mp_entry, [(f, r, np, base_port) for r in range(np)])
