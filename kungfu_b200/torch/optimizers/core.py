"""Shared machinery of the distributed optimizers.

Reference: `_KungFuAlgorithm` strategy objects wrapping a user optimizer (srcs/python/kungfu/tensorflow/optimizers/
core.py:6-72) and the dynamic-subclass trick of the torch binding (torch/optimizers/sync_sgd.py:29-32): the
returned object *is an instance of the user's optimizer class* (so LR schedulers, `isinstance` checks and
`state_dict` keep working) with `step()` / `zero_grad()` overridden.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Tuple

import torch

from ...ops import collective as coll
from ...runtime import get_runtime


def fusable_kind(optimizer: torch.optim.Optimizer) -> Optional[str]:
    """Returns "sgd"/"adam" if our fused kernels implement exactly this optimizer, else None."""
    cls = optimizer.__class__
    base = [c.__name__ for c in cls.__mro__]
    for g in optimizer.param_groups:
        if g.get("maximize", False) or g.get("amsgrad", False) or g.get("capturable", False):
            return None
        if g.get("differentiable", False):
            return None
    if "SGD" in base:
        for g in optimizer.param_groups:
            if g.get("nesterov", False) and (g.get("momentum", 0) <= 0 or g.get("dampening", 0) != 0):
                return None
        return "sgd"
    if "AdamW" in base:
        for g in optimizer.param_groups:
            g["_adamw"] = True
        return "adam"
    if "Adam" in base:
        return "adam"
    return None


def named_trainable(optimizer: torch.optim.Optimizer, named_parameters) -> List[Tuple[str, torch.nn.Parameter]]:
    """Stable (name, param) list covering every trainable parameter of the optimizer."""
    ids = set()
    for g in optimizer.param_groups:
        for p in g["params"]:
            ids.add(id(p))
    out = []
    if named_parameters is not None:
        for n, p in named_parameters:
            if id(p) in ids and p.requires_grad:
                out.append((n, p))
    if len(out) == 0:
        k = 0
        for gi, g in enumerate(optimizer.param_groups):
            for pi, p in enumerate(g["params"]):
                if p.requires_grad:
                    out.append((f"group{gi}.param{pi}", p))
                    k += 1
    return out


def all_cuda(params: Iterable[torch.nn.Parameter]) -> bool:
    ps = list(params)
    return len(ps) > 0 and all(p.is_cuda for p in ps)


def wrap_optimizer(optimizer: torch.optim.Optimizer, mixin: type, init: Callable) -> torch.optim.Optimizer:
    """Create an instance of a dynamic subclass of `optimizer.__class__` that shares the optimizer's state."""
    if getattr(optimizer, "_kf_wrapped", False):
        raise ValueError("kungfu_b200: optimizer is already wrapped")
    cls = type(optimizer.__class__.__name__, (mixin, optimizer.__class__), {})
    obj = cls.__new__(cls)
    obj.__dict__.update(optimizer.__dict__)
    obj._kf_wrapped = True
    obj._kf_inner_cls = optimizer.__class__
    init(obj)
    return obj


class KungFuMixin:
    """Common helpers; concrete optimizers add `step`."""

    def _kf_setup(self, named_parameters, use_gpu_engine: bool) -> None:
        self._kf_rt = get_runtime()
        self._kf_named = named_trainable(self, named_parameters)
        self._kf_gpu = use_gpu_engine and all_cuda(p for _, p in self._kf_named)
        if self._kf_gpu and self._kf_rt.gpu.spans_hosts:
            # the fused engines address every peer's HBM directly, which only exists inside one NVLink domain; a job that
            # spans hosts takes the generic path (hierarchical all-reduce: kernels inside the host, control plane across)
            self._kf_gpu = False
        self._kf_engine = None
        self._kf_step_count = 0

    def _kf_membership_changed(self) -> bool:
        """True once per cluster version: at the first `step()` after the membership changed (and at the very first step).
        Survivors reach their next step and newcomers their first one after the same `sync_after_resize`, so every member of the
        new cluster sees the change at the same point and collectives keyed to it stay matched. The per-process step counter is
        re-agreed here (max over the workers): schedules that depend on it (AdaptiveSGD's switch) then hold for joiners too."""
        ver = getattr(self._kf_rt.control, "cluster_version", None)
        key = (ver() if ver else 0, self._kf_rt.size)
        if getattr(self, "_kf_membership", None) == key:
            return False
        self._kf_membership = key
        if self._kf_rt.size > 1:
            from ...python import all_reduce_int_max

            self._kf_step_count = all_reduce_int_max(self._kf_step_count)
        return True

    def _kf_inner_step(self, closure=None):
        return self._kf_inner_cls.step(self, closure)

    def zero_grad(self, set_to_none: bool = False):  # noqa: D401 - torch API
        eng = getattr(self, "_kf_engine", None)
        if eng is not None:
            eng.zero_grad()
            self._kf_begin_step()
            return
        return self._kf_inner_cls.zero_grad(self, set_to_none=set_to_none)

    def _kf_begin_step(self) -> None:
        """Hook for optimizers that can start work before backward (SMA: model average, PairAveraging: the pull)."""

    def no_sync(self):
        """Gradient accumulation: `with opt.no_sync(): loss.backward()` for all but the last micro-batch (the overlapped
        engines reduce — and on the fused path apply — a bucket as soon as backward has produced it)."""
        eng = getattr(self, "_kf_engine", None)
        if eng is not None:
            return eng.no_sync()
        import contextlib

        return contextlib.nullcontext()

    @property
    def engine(self):
        return getattr(self, "_kf_engine", None)

    def cluster_size(self) -> int:
        return self._kf_rt.size


def cpu_all_reduce_grads(named, op: str, average: bool) -> None:
    """Reference torch path: one named all-reduce per gradient (torch/optimizers/sync_sgd.py:12-22)."""
    rt = get_runtime()
    if rt.size == 1:
        return
    cuda = [p for _, p in named if p.grad is not None and p.grad.is_cuda]
    if cuda and rt.gpu.spans_hosts:
        # one fused message per dtype instead of one host round trip per tensor
        by_dtype = {}
        for p in cuda:
            by_dtype.setdefault(p.grad.dtype, []).append(p.grad)
        for grads in by_dtype.values():
            flat = torch._utils._flatten_dense_tensors(grads)
            coll.all_reduce_(flat, "avg" if average else op, "fused-grads")
            for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                g.copy_(f)
        named = [(n, p) for n, p in named if p.grad is not None and not p.grad.is_cuda]
    handles = []
    for name, p in named:
        if p.grad is None:
            continue
        handles.append(coll.all_reduce_async_(p.grad, name, "avg" if average else op))  # host tensors overlap too (FIFO executor)
    coll.wait_all_handles(handles)
