"""Standby worker: a pre-started, warmed-up process that becomes a real worker the moment the cluster grows.

`kungfu-run -w -standby N` keeps N of these alive next to the running workers. Each one pays the expensive part of a worker
start in advance — interpreter start, `import torch`, loading the native extensions, creating the CUDA context on the GPU slot
the launcher reserved for it — and then blocks on stdin. When a resize adds a worker on this host the launcher writes that
worker's environment (KUNGFU_SELF_SPEC, KUNGFU_INIT_PEERS, cluster version, ...) to the pipe; the standby adopts it and runs the
training script as `__main__`. Joining a 4 -> 8 resize then costs the protocol + state broadcast, not process start-up.
By default the script is even started BEFORE the activation and held at its first use of the control plane (`runtime._activation_gate`),
so that model construction and the host-to-device copy of the model are off the critical path too (KUNGFU_STANDBY_EAGER=0 disables).

The reference starts every added worker cold (srcs/go/kungfu/runner/watch.go:42-83 fork/exec on each Stage update); its resize
latency is dominated by interpreter + framework start, as was ours (9 s, profiles/configs_8gpu.log of round 1).

    python -m kungfu_b200.cmd.standby train.py --args...        (started by the launcher, not by hand)
"""
from __future__ import annotations

import os
import runpy
import sys


def _warm_up() -> None:
    import torch

    try:
        from .. import _kfrt  # noqa: F401  (dlopen now, not at the first collective)
    except ImportError:
        pass
    want = os.environ.get("KUNGFU_CUDA_VISIBLE_DEVICES", "-1").split(",")[0].strip()
    if torch.cuda.is_available() and want not in ("", "-1"):
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(want) if visible is None else ([v.strip() for v in visible.split(",")].index(want) if want in visible.split(",") else -1)
        if idx >= 0:
            try:
                from .. import _kfcuda  # noqa: F401

                torch.cuda.set_device(idx)
                torch.zeros(1, device="cuda")  # context + allocator
                torch.backends.cudnn.benchmark = True
                torch.cuda.synchronize()
            except Exception as e:  # a standby that cannot warm up is still a valid (cold) worker
                print(f"[standby] CUDA warm-up skipped: {e}", file=sys.stderr, flush=True)


def _read_identity() -> bool:
    """Block until the launcher writes this worker's environment (one KEY=VALUE per line, blank line = go). False: the pipe was closed
    without an activation — the job ended and this standby was never needed."""
    for line in sys.stdin:
        if line in ("\n", "\r\n"):
            os.environ.pop("KUNGFU_STANDBY", None)
            return True
        k, _, v = line.rstrip("\n").partition("=")
        if k:
            os.environ[k] = v
    return False


def main() -> int:
    if len(sys.argv) < 2:
        print("usage: python -m kungfu_b200.cmd.standby prog.py args...", file=sys.stderr)
        return 2
    _warm_up()
    print("[standby] ready", file=sys.stderr, flush=True)
    sys.argv = sys.argv[1:]
    if os.environ.get("KUNGFU_STANDBY_EAGER", "1") != "0":
        # Run the script NOW: its imports, model construction and the copy of the model to the GPU need no identity (the GPU slot is
        # already in the environment). The first call that does — anything touching the control plane, e.g. building the distributed
        # optimizer — parks in the gate below until the launcher activates this process. A joiner then enters the resize with its model
        # already on the device (4 -> 8 grow on 8 x B200: 4.2 s with lazily started standbys, most of it the joiners' script start).
        from .. import runtime as _rt

        def gate() -> None:
            if not _read_identity():
                sys.stdout.flush()
                sys.stderr.flush()
                os._exit(0)

        _rt._activation_gate = gate
        runpy.run_path(sys.argv[0], run_name="__main__")
        return 0
    if not _read_identity():
        return 0  # the launcher closed the pipe: job over, never needed
    runpy.run_path(sys.argv[0], run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
