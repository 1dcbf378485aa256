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
from ..python import (all_reduce_int_max, current_cluster_size, current_rank, detached, failure_pending, propose_new_size,
                      recover, resize)


@dataclass
class ElasticState:
    step: int = 0
    trained_samples: int = 0
    epoch: int = 0
    recovering: bool = False  # set between a lost peer and the re-sync of the survivors


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
        if getattr(eng, "_lost", None) is not None or getattr(state, "recovering", False):
            eng.rebuild_state_after_loss()  # survivors of a crash: put the old shards back together first
        eng.sync_state_from_root(0)  # re-shard fp32 master weights / optimizer state for the new world size
    state.recovering = False
    return state


def is_peer_failure(exc: BaseException) -> bool:
    """Does this exception look like "a collective broke because a peer went away"? (host plane: closed endpoint / refused
    connection / lost connection; GPU plane: a device-side wait ran into its time-out)."""
    if failure_pending():
        return True
    msg = str(exc)
    return isinstance(exc, RuntimeError) and any(k in msg for k in (
        "collective endpoint closed", "cannot connect to", "send to", "connection lost", "device-side barrier timed out", "a peer lost", "peer lost"))


def cuda_context_lost(exc: BaseException) -> bool:
    """A peer died while this worker's kernels had its memory mapped: the GPU contains the fault (`cudaErrorContained`, "Invalid access of
    peer GPU memory over nvlink") and CUDA refuses all further work in this process — nothing on the device can be saved."""
    msg = str(exc)
    return "Invalid access of peer GPU memory" in msg or "cudaErrorContained" in msg


class ShrinkOnFailure:
    """Shrink-and-continue for a training loop under `kungfu-run -w -shrink-on-failure`.

        guard = ShrinkOnFailure(state, model, optimizer)
        while state.step < max_step:
            if guard.run(train_one_step):      # False: a peer died during this step, the survivors are re-synced — redo it
                state.step += 1
            if guard.detached: break

    `run` executes the step; if it raises because a peer died, the survivors adopt the cluster the runner announced
    (`kungfu_b200.python.recover`), reconnect the GPU arenas among themselves, agree on progress and take rank 0's model /
    optimizer state (`sync_after_resize`), and `run` returns False so the caller repeats the step with the smaller cluster.
    Anything that does not look like a lost peer is re-raised. SURVEY.md §5.3; the reference restarts the job instead
    (srcs/go/kungfu/runner/monitored.go:48-73).

    GPU plane: measured on 2 x B200 — when a worker dies while a survivor's collective kernels have its HBM mapped (always the case
    inside a fused S-SGD step), the survivor's next access faults, the GPU contains the error (`cudaErrorContained`) and CUDA refuses
    all further work in that process. That cannot be continued in-process by anybody (NCCL has the same limit); `run` reports it as
    such. Shrink-and-continue therefore serves the host plane (CPU tensors, control-plane collectives, elastic data loading) and
    GPU jobs between steps; a GPU worker killed mid-step needs `-auto-recover`."""

    def __init__(self, state: ElasticState, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None,
                 timeout_s: float = 30.0, max_failures: int = 16, follow_config_server: bool = False, poll_every: int = 10) -> None:
        self.state, self.model, self.optimizer = state, model, optimizer
        self.timeout_s = float(timeout_s)
        self.max_failures = int(max_failures)
        self.failures: List[Tuple[int, int, float]] = []  # (old size, new size, seconds to be training again)
        self.regrown: List[Tuple[int, int, float]] = []  # (old size, new size, seconds) of resizes adopted from the config server
        self.detached = False
        # `kungfu-run -w -shrink-on-failure -refill`: after a loss the runner proposes a cluster with a replacement worker (taken from its
        # standby pool when it has one) to the config server; the survivors pick it up here, every `poll_every` steps, through the
        # ordinary resize protocol — all of them at the same step, because `state.step` is agreed.
        self.follow_config_server = bool(follow_config_server)
        self.poll_every = max(1, int(poll_every))

    def _follow(self) -> None:
        if not self.follow_config_server or (self.state.step + 1) % self.poll_every != 0:
            return
        t0 = time.time()
        old = current_cluster_size()
        changed, det = resize(None)
        if det:
            self.detached = True
        elif changed:
            self.state.step += 1  # the step that just ran counts; newcomers adopt max(step)
            sync_after_resize(self.state, self.model, self.optimizer)
            self.state.step -= 1
            self.regrown.append((old, current_cluster_size(), time.time() - t0))

    def run(self, step_fn, *args, **kwargs) -> bool:
        try:
            step_fn(*args, **kwargs)
            self._follow()
            return True
        except Exception as e:  # noqa: BLE001 - classified below
            if cuda_context_lost(e):
                raise RuntimeError(
                    "kungfu_b200: a peer was lost while its HBM was mapped into this worker's kernels; the GPU contained the NVLink fault and "
                    "this process' CUDA context is gone (CUDA requires a relaunch). Shrink-and-continue covers the host plane and GPU "
                    "workers whose kernels were not touching the dead peer; for this case run under `kungfu-run -auto-recover` and "
                    "resume from the last checkpoint.") from e
            if not is_peer_failure(e) or len(self.failures) >= self.max_failures:
                raise
            t0 = time.time()
            old = current_cluster_size()
            try:
                changed, det = recover(self.timeout_s)
            except Exception:
                raise e  # nobody announced a membership change: it was not a lost peer after all
            if det:
                self.detached = True
                return False
            if changed:
                self.state.recovering = True  # type: ignore[attr-defined]
                sync_after_resize(self.state, self.model, self.optimizer)
            self.failures.append((old, current_cluster_size(), time.time() - t0))
            return False


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
