"""Policy framework: user policies observe and adapt training (batch size, cluster size) between steps / epochs.

Reference: srcs/python/kungfu/tensorflow/policy/{base_policy.py, policy_hook.py} — `BasePolicy{before_train,
before_epoch, before_step, after_step, after_epoch, after_train}`; `PolicyHook(policies, epoch_size, epoch_num,
init_batch_size)` maintains `kungfu_trained_samples`, `kungfu_total_samples`, `kungfu_batch_size`, drives the
policies and stops on total samples or `detached()`; test helper `ScheduledBatchSizePolicy`
(tests/python/unit/scheduled_batch_size_policy.py).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

from .. import variables as kv
from ..python import current_cluster_size, detached

_PUBLISHED = (kv.GraphKeys.BATCH_SIZE, kv.GraphKeys.TRAINED_SAMPLES, kv.GraphKeys.TOTAL_SAMPLES)


class BasePolicy:
    def before_train(self, vars: Dict) -> None:
        pass

    def before_epoch(self, vars: Dict) -> None:
        pass

    def before_step(self, vars: Dict) -> None:
        pass

    def after_step(self, vars: Dict) -> None:
        pass

    def after_epoch(self, vars: Dict) -> None:
        pass

    def after_train(self, vars: Dict) -> None:
        pass


class PolicyHook:
    """Drives policies around a user training loop:

        hook = PolicyHook([policy], epoch_size, epoch_num, init_batch_size); hook.begin()
        while hook.before_step(): train_one_step(hook.batch_size); hook.after_step()
        hook.end()
    """

    def __init__(self, policies: Sequence[BasePolicy], epoch_size: int, epoch_num: int, init_batch_size: int) -> None:
        self.policies: List[BasePolicy] = list(policies)
        self.epoch_size = int(epoch_size)
        self.epoch_num = int(epoch_num)
        self.vars: Dict = {
            "kungfu_trained_samples": 0,
            "kungfu_total_samples": int(epoch_size) * int(epoch_num),
            "kungfu_batch_size": int(init_batch_size),
            "kungfu_gradient_noise_scale": 0.0,
            "epoch": 0,
            "step": 0,
        }
        self._in_epoch = False
        self._stopped = False

    @property
    def batch_size(self) -> int:
        return int(self.vars["kungfu_batch_size"])

    def set_batch_size(self, bs: int) -> None:
        self.vars["kungfu_batch_size"] = int(bs)
        self._publish()

    def _publish(self) -> None:
        """Mirror the counters into the global-variable registry (reference: PolicyHook owns `kungfu_batch_size`,
        `kungfu_trained_samples`, `kungfu_total_samples`), so that `variables.eval_batch_size()` works anywhere in the program."""
        import torch

        for name in _PUBLISHED:
            kv.get_or_create_global_variable(name, (), torch.int64).fill_(int(self.vars[name]))

    def begin(self) -> None:
        self._publish()
        for p in self.policies:
            p.before_train(self.vars)
        self._publish()

    def before_step(self) -> bool:
        if self._stopped or detached() or self.vars["kungfu_trained_samples"] >= self.vars["kungfu_total_samples"]:
            return False
        if not self._in_epoch:
            for p in self.policies:
                p.before_epoch(self.vars)
            self._in_epoch = True
        for p in self.policies:
            p.before_step(self.vars)
        self._publish()
        return True

    def after_step(self) -> None:
        self.vars["kungfu_trained_samples"] += self.batch_size * current_cluster_size()
        self.vars["step"] += 1
        for p in self.policies:
            p.after_step(self.vars)
        if self.vars["kungfu_trained_samples"] >= (self.vars["epoch"] + 1) * self.epoch_size:
            for p in self.policies:
                p.after_epoch(self.vars)
            self.vars["epoch"] += 1
            self._in_epoch = False
        gns = kv.get_global_variable(kv.GraphKeys.GRADIENT_NOISE_SCALE)
        if gns is not None and not gns.is_cuda:  # a device-resident estimate is read on demand, not every step
            self.vars["kungfu_gradient_noise_scale"] = float(gns.item())
        self._publish()

    def request_stop(self) -> None:
        self._stopped = True

    def end(self) -> None:
        for p in self.policies:
            p.after_train(self.vars)


class ScheduledBatchSizePolicy(BasePolicy):
    """Sets the batch size per epoch from a schedule {epoch: batch_size}."""

    def __init__(self, schedule: Dict[int, int]) -> None:
        self.schedule = dict(schedule)

    def before_epoch(self, vars: Dict) -> None:
        e = vars["epoch"]
        if e in self.schedule:
            vars["kungfu_batch_size"] = int(self.schedule[e])
