"""Step-level logging hooks for plain training loops (reference: benchmarks/scaling/debug_hooks.py:22-110 — `LogPerfHook`,
`LogStepHook` as tf.train.SessionRunHooks). Call `after_run()` once per step and `end()` when the loop finishes."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from kungfu_b200.utils import _since_job_start, show_duration  # noqa: E402


def _cluster_size() -> int:
    if os.getenv("KUNGFU_SELF_SPEC") or os.getenv("WORLD_SIZE"):
        from kungfu_b200.python import current_cluster_size

        return current_cluster_size()
    return 1


class LogPerfHook:
    """Per-step duration and samples/sec; `end()` prints `RESULT: <samples/sec>, batch size, cluster size` over the steps after
    the warm-up (the last `drop_last` steps are dropped: a resize usually ends the phase)."""

    def __init__(self, batch_size: int, warmup_steps: int = 5, drop_last: int = 1, quiet: bool = False):
        self._batch_size, self._warmup_steps, self._drop_last, self._quiet = batch_size, warmup_steps, drop_last, quiet
        self._step = 0
        self._durations = []
        self._t_last = time.time()

    def before_run(self):
        if self._step == 0:
            print("_since_job_start: %s" % show_duration(_since_job_start()), flush=True)
            self._t_last = time.time()

    def after_run(self):
        self._step += 1
        t1 = time.time()
        dur = t1 - self._t_last
        if not self._quiet:
            print("local step %d, took %.3fs, %.2f samples / sec" % (self._step, dur, self._batch_size / max(dur, 1e-9)), flush=True)
        if self._step > self._warmup_steps:
            self._durations.append(dur)
        self._t_last = t1

    def end(self):
        ds = np.array(self._durations[:max(0, len(self._durations) - self._drop_last)] or self._durations or [float("nan")])
        mean = float(ds.mean())
        print("mean_duration: %.3fs" % mean, flush=True)
        rate = self._batch_size / mean if mean == mean and mean > 0 else 0.0
        print("RESULT: %.2f samples / sec, batch size: %d, cluster size %d" % (rate, self._batch_size, _cluster_size()), flush=True)
        return rate


class LogStepHook:
    def __init__(self):
        self._step = 0

    def begin(self):
        print("%s::%s %d steps" % ("LogStepHook", "begin", self._step), flush=True)

    def after_run(self):
        self._step += 1

    def end(self):
        print("%s::%s after %d steps" % ("LogStepHook", "end", self._step), flush=True)
