"""Elastic training loop hooks for PyTorch.

Reference: `KungFuElasticTrainHook(schedule, max_step, model_dir)` (srcs/python/kungfu/tensorflow/hooks/elastic.py:14-87:
every step evaluate the step-based schedule -> propose_new_size -> resize_cluster_from_url; on change all-reduce-max the
step then broadcast the variables; stop when detached; `save()` dumps variables) and `experimental.hook.ElasticHook
(local_batch_size, epochs, epoch_size)` (experimental/hook/elastic.py:11-116: tracks trained samples, syncs the offset by
all-reduce-max, `ResizeProfiler` timings). State transfer on resize is in memory: survivors keep theirs, newcomers
receive rank 0's by broadcast (§5.4 of SURVEY.md).
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from ..initializer import BroadcastGlobalVariablesOp
from ..ops.adapt import StepBasedSchedule
from ..python import all_reduce_int_max, current_cluster_size, current_rank, detached, propose_new_size, resize


@dataclass
class ElasticState:
    step: int = 0
    trained_samples: int = 0
    epoch: int = 0


class ResizeProfiler:
    """Records (old size, new size, seconds) of every resize (reference: experimental/hook/elastic.py:11-47)."""

    def __init__(self) -> None:
        self.records: List[Tuple[int, int, float]] = []
        self._t0 = 0.0
        self._old = 0

    def begin(self) -> None:
        self._t0 = time.time()
        self._old = current_cluster_size()

    def end(self) -> None:
        self.records.append((self._old, current_cluster_size(), time.time() - self._t0))

    def report(self) -> str:
        return "\n".join(f"resize {a} -> {b} took {d * 1e3:.1f}ms" for a, b, d in self.records)


def sync_after_resize(state: ElasticState, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None) -> ElasticState:
    """What every member does right after `changed`: agree on progress (max), adopt rank 0's model/optimizer state."""
    from ..runtime import get_runtime

    rt = get_runtime()
    if rt.has_gpu() and (rt._gpu is not None or (optimizer is not None and getattr(optimizer, "engine", None) is not None)):
        rt.gpu.reconnect_all()  # re-discover peers, rebuild VMM mappings + NVLS multicast for the new membership
    state.step = all_reduce_int_max(state.step)
    state.trained_samples = all_reduce_int_max(state.trained_samples)
    state.epoch = all_reduce_int_max(state.epoch)
    eng = getattr(optimizer, "engine", None) if optimizer is not None else None
    BroadcastGlobalVariablesOp(model, optimizer if eng is None else None, 0)
    if eng is not None:
        eng.sync_state_from_root(0)  # re-shard fp32 master weights / optimizer state for the new world size
    return state


class KungFuElasticTrainHook:
    """Step-schedule driven elasticity: call `after_step()` once per training step; stop when it returns False."""

    def __init__(self, schedule: str, max_step: int, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None,
                 model_dir: Optional[str] = None) -> None:
        self.schedule = StepBasedSchedule(schedule)
        self.max_step = int(max_step)
        self.model, self.optimizer = model, optimizer
        self.model_dir = model_dir
        self.state = ElasticState()
        self.profiler = ResizeProfiler()
        self.need_sync = True  # a fresh (late-joining) worker must sync first

    def begin(self) -> None:
        if self.need_sync:
            sync_after_resize(self.state, self.model, self.optimizer)
            self.need_sync = False

    def after_step(self) -> bool:
        self.state.step += 1
        if self.state.step >= self.max_step:
            return False
        want = self.schedule(self.state.step)
        if want != current_cluster_size():
            self.profiler.begin()
            if current_rank() == 0:
                try:
                    propose_new_size(want)
                except Exception:
                    pass
            changed, det = resize(None)
            self.profiler.end()
            if det:
                return False
            if changed:
                sync_after_resize(self.state, self.model, self.optimizer)
        return not detached()

    def save(self, idx: int = 0) -> Optional[str]:
        if self.model_dir is None:
            return None
        os.makedirs(self.model_dir, exist_ok=True)
        path = os.path.join(self.model_dir, f"variables-{idx}.pt")
        torch.save({k: v.detach().cpu() for k, v in self.model.state_dict().items()}, path)
        return path


class ElasticHook:
    """Sample-count driven elastic loop (reference: experimental/hook/elastic.py:49-116)."""

    def __init__(self, local_batch_size: int, epochs: int, epoch_size: int, model: torch.nn.Module,
                 optimizer: Optional[torch.optim.Optimizer] = None) -> None:
        self.local_batch_size = int(local_batch_size)
        self.total_samples = int(epochs) * int(epoch_size)
        self.model, self.optimizer = model, optimizer
        self.state = ElasticState()
        self.profiler = ResizeProfiler()

    def begin(self) -> None:
        sync_after_resize(self.state, self.model, self.optimizer)

    def after_step(self, new_size: Optional[int] = None) -> bool:
        """Account the step; optionally resize to `new_size`. Returns False when training should stop."""
        self.state.step += 1
        self.state.trained_samples += self.local_batch_size * current_cluster_size()
        if new_size is not None and new_size != current_cluster_size():
            self.profiler.begin()
            changed, det = resize(new_size)
            self.profiler.end()
            if det:
                return False
            if changed:
                sync_after_resize(self.state, self.model, self.optimizer)
        return self.state.trained_samples < self.total_samples and not detached()
