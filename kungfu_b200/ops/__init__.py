"""Operator-level API (reference: srcs/python/kungfu/tensorflow/ops/__init__.py:29-83).

GPU tensors run on our sm_100a kernels over peer-mapped HBM; CPU tensors on the native graph-collective runtime.
"""
from typing import List, Sequence

import torch

from .adapt import StepBasedSchedule, resize, resize_cluster, resize_cluster_from_url, set_tree, step_based_schedule
from .collective import (Handle, all_gather, all_gather_transform, all_reduce, all_reduce_, all_reduce_async_, barrier, broadcast_, consensus, subset_all_reduce,
                         wait_all_handles, wait_handle)
from .monitor import NoiseScale, egress_rates, global_noise_scale
from .p2p import (AsyncModelAveraging, ModelAveraging, RandomSelector, RoundRobinSelector, request_variable, request_variable_with_template,
                  save_variable, save_variables)
from .state import Counter, ExponentialMovingAverage, counter, exponential_moving_average
from .topology import (cluster_size, get_neighbour_mask, get_peer_latencies, global_minimum_spanning_tree,
                       minimum_spanning_tree_father, peer_info, rank, round_robin)


def broadcast(t: torch.Tensor, root: int = 0) -> torch.Tensor:
    return broadcast_(t.clone(), root)


def group_all_reduce(ts: Sequence[torch.Tensor], op="sum") -> List[torch.Tensor]:
    """All-reduce a list of tensors (None entries pass through, as in the reference's group_all_reduce)."""
    return [None if t is None else all_reduce(t, op) for t in ts]


def monitored_all_reduce(t: torch.Tensor, tree=None, op="sum") -> torch.Tensor:
    """All-reduce that also feeds the strategy throughput statistics (calc_stats / check_interference)."""
    from ..runtime import get_runtime

    rt = get_runtime()
    f = getattr(rt.control, "monitored_all_reduce_cpu_", None)
    if t.is_cuda or f is None or rt.size == 1:
        return all_reduce(t, op)
    out = t.detach().clone().contiguous()
    f(out, None if tree is None else [int(x) for x in (tree.tolist() if isinstance(tree, torch.Tensor) else tree)], op)
    return out


def group_nccl_all_reduce(ts: Sequence[torch.Tensor], op="sum") -> List[torch.Tensor]:
    """All-reduce a list of device tensors concurrently. The reference schedules one NCCL all-reduce per tensor in a
    negotiated order (ops/collective.py:96-106); here each tensor goes to the sm_100a kernels on the comm stream (their order is
    the list order on every rank), host tensors to the native runtime. None entries pass through."""
    outs = [None if t is None else t.detach().clone().contiguous() for t in ts]
    from .collective import all_reduce_group_async_

    all_reduce_group_async_([t for t in outs if t is not None], op).wait()
    return outs


def group_hierarchical_nccl_all_reduce(ts: Sequence[torch.Tensor], op="sum") -> List[torch.Tensor]:
    """Reference: local reduce -> cross-host all-reduce -> local broadcast, per tensor (ops/collective.py:109-135).

    Host tensors run exactly those three steps on the native runtime (`Session::local_reduce` to the first rank of every host,
    `cross_all_reduce` among the hosts' first ranks, `local_broadcast` back) — on one host that degenerates to reduce + broadcast inside
    the host. Device tensors go through `all_reduce`'s dispatcher, which IS hierarchical when the job spans hosts (kernel all-reduce
    inside the NVLink group, control plane between the groups' first ranks, broadcast kernel) and a single flat kernel inside one NVLink
    domain, where a three-step schedule would only add two passes over the data."""
    from ..runtime import get_runtime

    ctl = get_runtime().control
    three_step = all(hasattr(ctl, n) for n in ("local_reduce_cpu", "cross_all_reduce_cpu_", "local_broadcast_cpu_"))
    outs: List = []
    device: List[torch.Tensor] = []
    for t in ts:
        if t is None:
            outs.append(None)
        elif t.is_cuda or not three_step or get_runtime().size == 1:
            c = t.detach().clone().contiguous()
            outs.append(c)
            if t.is_cuda or get_runtime().size > 1:
                device.append(c)
        else:
            part = ctl.local_reduce_cpu(t.detach().contiguous(), op)  # the host's first rank holds the host's partial result
            ctl.cross_all_reduce_cpu_(part, op)  # ... combines it with the other hosts' (a no-op exchange for everybody else)
            ctl.local_broadcast_cpu_(part)  # ... and hands the total back to its host
            outs.append(part)
    if device:
        from .collective import all_reduce_group_async_

        all_reduce_group_async_(device, op).wait()
    return outs


def calc_stats():
    """Communication-strategy throughput statistics (reference: ops/adapt.py:60-62)."""
    from ..python import calc_stats as _calc

    return _calc()


def save_model(variables: Sequence[torch.Tensor], name: str = "kungfu::fused_model") -> None:
    """Publish the fused model (all variables flattened into one blob) for peers to pull (reference: ops/local.py:27-33,
    the store behind PairAveraging with fuse_requests=True)."""
    save_variable(fuse([v.detach() for v in variables]), name)


def request_model(target: int, variables: Sequence[torch.Tensor], name: str = "kungfu::fused_model"):
    """Pull the fused model published by `target` with `save_model`; returns (ok, list of tensors shaped like `variables`)."""
    template = fuse([v.detach() for v in variables])
    ok, flat = request_variable_with_template(target, template, name)
    return ok, (defuse(flat, [tuple(v.shape) for v in variables]) if ok else [])


def fuse(ts: Sequence[torch.Tensor]) -> torch.Tensor:
    """Flatten + concatenate (reference: ops/__init__.py:29-33)."""
    return torch.cat([t.reshape(-1) for t in ts])


def defuse(y: torch.Tensor, shapes) -> List[torch.Tensor]:
    """Inverse of `fuse` (reference: ops/__init__.py:36-46)."""
    out, off = [], 0
    for s in shapes:
        n = 1
        for d in s:
            n *= int(d)
        out.append(y[off:off + n].view(tuple(s)))
        off += n
    if off != y.numel():
        raise ValueError("defuse: shapes do not cover the fused tensor")
    return out
