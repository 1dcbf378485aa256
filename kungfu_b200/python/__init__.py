"""`kungfu.python` parity: process identity and online control.

Reference: srcs/python/kungfu/python/__init__.py:17-155 (ctypes into libkungfu); here a thin layer over the
`Runtime` singleton.
"""
from __future__ import annotations

from ..runtime import get_runtime, reset_runtime


def uid() -> int:
    rt = get_runtime()
    f = getattr(rt.control, "uid", None)
    return int(f()) if f else 0


def detached() -> bool:
    rt = get_runtime()
    f = getattr(rt.control, "detached", None)
    return bool(f()) if f else False


def current_rank() -> int:
    return get_runtime().rank


def current_local_rank() -> int:
    return get_runtime().local_rank


def current_cluster_size() -> int:
    return get_runtime().size


def current_local_size() -> int:
    return get_runtime().local_size


def run_barrier() -> None:
    from ..ops.collective import barrier

    barrier()


def _get_cuda_index() -> int:
    return get_runtime().cuda_index()


def _get_other_ranks():
    rt = get_runtime()
    return [r for r in range(rt.size) if r != rt.rank]


def show_cuda_version() -> None:
    import torch

    print("CUDA (torch build):", torch.version.cuda)


def show_nccl_version() -> None:
    import torch

    try:
        print("NCCL:", ".".join(str(x) for x in torch.cuda.nccl.version()))
    except Exception as e:  # pragma: no cover
        print("NCCL: unavailable", e)


def init_from_env() -> None:
    get_runtime()


def finalize() -> None:
    reset_runtime()
