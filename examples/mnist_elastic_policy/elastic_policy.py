"""A policy that resizes the cluster on a step schedule and keeps the replicas consistent across the change.

Reference: examples/mnist_elastic_policy/elastic_policy.py:9-47 (`ScheduledElasticPolicy(BasePolicy)`: `after_step` looks the global
step up in `{step: size}`, runs the resize op, and `before_step` broadcasts the variables when a change was seen). Here the policy works
on the `vars` dict the `PolicyHook` hands to every callback, and the state it re-synchronises is (step counters, model, optimizer).
"""
import kungfu_b200.python as kfp
from kungfu_b200.elastic import BasePolicy, ElasticState, sync_after_resize


class ScheduledElasticPolicy(BasePolicy):
    def __init__(self, step_schedule, model, optimizer):
        self._schedule = dict(step_schedule)
        self._model, self._optimizer = model, optimizer
        self._need_sync = True  # a worker that joins late starts here: it adopts the survivors' progress before its first step
        self.resizes = []

    def before_step(self, vars):
        if self._need_sync:
            state = ElasticState(step=vars["step"], trained_samples=vars["kungfu_trained_samples"], epoch=vars["epoch"])
            state = sync_after_resize(state, self._model, self._optimizer)
            vars["step"], vars["kungfu_trained_samples"], vars["epoch"] = state.step, state.trained_samples, state.epoch
            self._need_sync = False

    def after_step(self, vars):
        new_size = self._schedule.get(vars["step"])
        if new_size is None or new_size == kfp.current_cluster_size():
            return
        old = kfp.current_cluster_size()
        changed, detached = kfp.resize(new_size)
        if changed:
            self.resizes.append((vars["step"], old, new_size))
            self._need_sync = not detached  # a detached worker stops: PolicyHook.before_step() returns False for it
