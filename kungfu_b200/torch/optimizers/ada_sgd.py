"""AdaptiveSGDOptimizer: synchronous model averaging until `change_step`, synchronous SGD afterwards.

Reference (TF only): srcs/python/kungfu/tensorflow/optimizers/ada_sgd.py:12-83 (`tf.cond` on the global step; the
`AdaSGDHook` re-broadcasts the variables at the switch so that all replicas continue from the same model).
"""
from __future__ import annotations

import torch

from ...ops import collective as coll
from .core import KungFuMixin, cpu_all_reduce_grads, wrap_optimizer


class _AdaptiveSGD(KungFuMixin):
    def _kf_init_ada(self, named_parameters, change_step, alpha):
        self._kf_setup(named_parameters, False)
        self._kf_change_step = int(change_step)
        self._kf_alpha = float(alpha)
        self._kf_switched = False

    def _kf_average_model(self):
        n = self._kf_rt.size
        if n == 1:
            return
        for name, p in self._kf_named:
            avg = p.data.detach().clone()
            coll.all_reduce_(avg, "avg", "ada::" + name)
            p.data.mul_(1.0 - self._kf_alpha).add_(avg, alpha=self._kf_alpha)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._kf_membership_changed()  # joiners adopt the survivors' step count: everybody is in the same phase
        if self._kf_step_count < self._kf_change_step:
            self._kf_average_model()
        else:
            if not self._kf_switched and self._kf_step_count == self._kf_change_step:
                # AdaSGDHook: one broadcast so that S-SGD starts from a single model
                for name, p in self._kf_named:
                    coll.broadcast_(p.data, 0, "ada::switch::" + name)
            self._kf_switched = True  # a worker that joins after the switch inherited the model through sync_after_resize
            cpu_all_reduce_grads(self._kf_named, "sum", True)
        self._kf_inner_step()
        self._kf_step_count += 1
        return loss


def AdaptiveSGDOptimizer(optimizer, named_parameters=None, change_step: int = 0, alpha: float = 0.1):
    return wrap_optimizer(optimizer, _AdaptiveSGD, lambda o: o._kf_init_ada(named_parameters, change_step, alpha))


class AdaSGDHook:
    """Training-loop hook of the reference (`AdaSGDHook(change_step)`, ada_sgd.py:68-83): at `change_step` every variable is
    re-broadcast from rank 0. `AdaptiveSGDOptimizer` above does this by itself; the hook exists for loops that drive the
    switch themselves (e.g. with a plain SMA optimizer before and an S-SGD optimizer after)."""

    def __init__(self, change_step: int, model: torch.nn.Module = None, tensors=None) -> None:
        self.change_step = int(change_step)
        self._tensors = list(tensors) if tensors is not None else None
        self._model = model
        self.fired = False

    def _targets(self):
        if self._tensors is not None:
            return self._tensors
        if self._model is not None:
            return [v for v in self._model.state_dict().values() if isinstance(v, torch.Tensor)]
        return []

    def after_run(self, global_step: int) -> bool:
        """Call once per step with the step that just finished; returns True on the step where the broadcast happened."""
        if self.fired or int(global_step) != self.change_step:
            return False
        for i, t in enumerate(self._targets()):
            coll.broadcast_(t, 0, f"ada::hook::{i}")
        self.fired = True
        return True

    after_step = after_run
