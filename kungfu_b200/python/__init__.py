"""`kungfu.python` parity: process identity and online control.

Reference: srcs/python/kungfu/python/__init__.py:17-155 (ctypes into libkungfu / libkungfu_python): `uid, detached,
current_rank, current_local_rank, current_cluster_size, current_local_size, run_barrier, propose_new_size,
resize(n=None) -> (changed, detached), all_reduce_int_max, check_interference, calc_stats, log_stats,
print_strategy_stats, show_cuda_version, show_nccl_version, _get_cuda_index, _get_other_ranks`.
Here: a thin layer over the `Runtime` singleton (native C++ peer underneath).
"""
from __future__ import annotations

from typing import Optional, Tuple

from ..runtime import get_runtime, reset_runtime


def _ctl():
    return get_runtime().control


def uid() -> int:
    """64-bit identity of this worker process: (ipv4, port, initial cluster version); 0 in single-process mode (python/__init__.py:50-53)."""
    f = getattr(_ctl(), "uid", None)
    return int(f()) if f else 0


def detached() -> bool:
    """True once a resize removed this worker from the cluster: it must stop training and exit (python/__init__.py:55-58)."""
    f = getattr(_ctl(), "detached", None)
    return bool(f()) if f else False


def current_rank() -> int:
    """Index of this worker in the current peer list (changes with resizes)."""
    return get_runtime().rank


def current_local_rank() -> int:
    """Index among the workers that share this worker's host."""
    return get_runtime().local_rank


def current_cluster_size() -> int:
    """Number of workers in the current cluster version."""
    return get_runtime().size


def current_local_size() -> int:
    """Number of workers on this worker's host."""
    return get_runtime().local_size


def current_cluster_version() -> int:
    """Membership epoch: starts at KUNGFU_INIT_CLUSTER_VERSION, +1 per accepted resize; also the token that fences stale connections."""
    f = getattr(_ctl(), "cluster_version", None)
    return int(f()) if f else 0


def run_barrier() -> None:
    """Block until every worker arrived (host barrier, plus a device flag barrier when a GPU plane exists)."""
    from ..ops.collective import barrier

    barrier()


def propose_new_size(new_size: int) -> None:
    """Rank 0 asks the config server for a cluster of `new_size` workers (takes effect at the next `resize()`)."""
    f = getattr(_ctl(), "propose_new_size", None)
    if f is None:
        raise RuntimeError("kungfu_b200: elastic control needs the native runtime (kungfu-run -w)")
    f(int(new_size))


def resize(n: Optional[int] = None) -> Tuple[bool, bool]:
    """Resize the cluster to `n` workers (or to whatever the config server says when n is None).

    Returns (changed, detached). After `changed` the caller re-syncs its step (all-reduce max) and state (broadcast from
    rank 0); a `detached` worker must stop. Reference: kungfu.python.resize (python/__init__.py:96-116) ->
    Peer::ResizeCluster (srcs/go/kungfu/peer/peer.go:227-263). GPU side: every symmetric arena re-discovers its peers
    and rebuilds the mappings (PeerMesh::connect), the counterpart of ResetNcclHelper.
    """
    rt = get_runtime()
    f = getattr(rt.control, "resize", None)
    if f is None:
        return False, False
    if rt._gpu is not None:
        import torch

        torch.cuda.synchronize()
        for e in rt._gpu.engines:
            e.prepare_resize()
    from ..utils.trace import trace_scope

    with trace_scope("resize"):
        changed, det = f(n)
    if changed and not det and rt._gpu is not None:
        # peer mappings are rebuilt collectively by ALL members of the new cluster in `sync_after_resize`
        for e in rt._gpu.engines:
            e.finish_resize()
    return changed, det


def recover(timeout_s: float = 30.0) -> Tuple[bool, bool]:
    """Shrink-and-continue: call after a collective raised because a worker died under `kungfu-run -w -shrink-on-failure`.

    The runner that saw the process die announces the cluster without it; this call adopts that cluster (no consensus with
    the dead peer), rebuilds the session among the survivors and returns (changed, detached) like `resize()`. The caller then
    re-syncs step and state exactly as after a resize (`elastic.sync_after_resize`): the interrupted step may have been applied
    on some survivors and not on others, rank 0's copy becomes everybody's. The reference has no counterpart: its failure
    detector restarts every worker from the last checkpoint (srcs/go/kungfu/runner/monitored.go:48-73).
    """
    rt = get_runtime()
    f = getattr(rt.control, "recover", None)
    if f is None:
        raise RuntimeError("kungfu_b200: recover() needs the native runtime (kungfu-run -w -shrink-on-failure)")
    if rt._gpu is not None:
        import torch

        try:
            torch.cuda.synchronize()  # kernels spinning on the dead rank's flags run into their device time-out
        except Exception:
            pass
        for e in rt._gpu.engines:
            e.prepare_recover()
    from ..utils.trace import trace_scope

    with trace_scope("recover"):
        changed, det = f(timeout_s)
    if changed and not det and rt._gpu is not None:
        for e in rt._gpu.engines:
            e.finish_resize()
    return changed, det


def failure_pending() -> bool:
    """True once a runner announced a lost worker that `recover()` has not adopted yet (collectives fail until then)."""
    f = getattr(_ctl(), "failure_pending", None)
    return bool(f()) if f else False


def all_reduce_int_max(x: int) -> int:
    """Max of a Python int over all workers — how newcomers adopt the survivors' step counter (python/__init__.py:152-155)."""
    f = getattr(_ctl(), "all_reduce_int_max", None)
    return int(f(int(x))) if f else int(x)


def check_interference() -> bool:
    """Majority vote "collective throughput dropped below 80 % of the reference window" (session/adaptiveStrategies.go:61-127)."""
    f = getattr(_ctl(), "check_interference", None)
    return bool(f()) if f else False


def calc_stats() -> None:
    """Recompute the throughput statistics of the active communication strategies."""
    f = getattr(_ctl(), "calc_stats", None)
    if f:
        f()


def log_stats() -> None:
    """Snapshot the strategy statistics into the runtime log."""
    f = getattr(_ctl(), "log_stats", None)
    if f:
        f()


def print_strategy_stats() -> None:
    """Print per-strategy accumulated bytes / time / throughput."""
    f = getattr(_ctl(), "print_strategy_stats", None)
    if f:
        f()


def _get_cuda_index() -> int:
    """Logical CUDA device of this worker (KUNGFU_CUDA_VISIBLE_DEVICES mapped through CUDA_VISIBLE_DEVICES), -1 without a GPU."""
    return get_runtime().cuda_index()


def _get_other_ranks():
    """Every rank but this one (peer selection of PairAveraging)."""
    rt = get_runtime()
    return [r for r in range(rt.size) if r != rt.rank]


def show_cuda_version() -> None:
    """Print the CUDA version torch was built with (reference prints the runtime libcudart version)."""
    import torch

    print("CUDA (torch build):", torch.version.cuda)


def show_nccl_version() -> None:
    """Print the NCCL version bundled with torch (baseline arm only: our collectives do not use NCCL)."""
    import torch

    try:
        print("NCCL:", ".".join(str(x) for x in torch.cuda.nccl.version()))
    except Exception as e:  # pragma: no cover
        print("NCCL: unavailable", e)


def init_from_env() -> None:
    get_runtime()


def finalize() -> None:
    reset_runtime()


def _resize_from_url():
    """Reference name of `resize()` without an argument (python/__init__.py:135-140): follow the config server."""
    return resize(None)
