"""Shrink-and-continue: a worker crashes in the middle of training, the others carry on with a smaller cluster.

The reference's failure handling restarts every worker from the last checkpoint (runner/monitored.go:48-73, the
`torch_recovery.py` example next to this file does the same here). Under `kungfu-run -w -shrink-on-failure` nothing is restarted:
the runner that sees a worker die announces the cluster without it, the survivors' collectives abort, `ShrinkOnFailure` adopts
the new cluster, takes rank 0's model / optimizer state and repeats the interrupted step.

    kungfu-run -w -shrink-on-failure -np 4 -builtin-config-port 9100 python examples/failure_recovery/torch_shrink.py --crash-rank 2 --crash-step 30
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "../..")))  # run from a checkout without installing

import argparse

import torch
import torch.nn.functional as F

import kungfu_b200.torch as kf
from kungfu_b200.elastic import ElasticState, ShrinkOnFailure, sync_after_resize
from kungfu_b200.models.slp import SLP, synthetic_mnist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--batch", type=int, default=60)
    ap.add_argument("--crash-rank", type=int, default=-1, help="this rank kills itself (fault injection)")
    ap.add_argument("--crash-step", type=int, default=30)
    a = ap.parse_args()
    model = SLP()
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9),
                                                named_parameters=model.named_parameters())
    state = ElasticState()
    sync_after_resize(state, model, opt)  # everybody starts from rank 0's initial model
    guard = ShrinkOnFailure(state, model, opt)
    x, y = synthetic_mnist(6000)
    loss_value = [0.0]

    def one_step():
        n, r = kf.current_cluster_size(), kf.current_rank()
        if r == a.crash_rank and state.step == a.crash_step and not guard.failures:
            os._exit(1)  # no goodbye: the runner notices the exit code
        s = (state.step * a.batch * n) % (6000 - a.batch * n)
        xb, yb = x[s + r * a.batch:s + (r + 1) * a.batch], y[s + r * a.batch:s + (r + 1) * a.batch]
        opt.zero_grad()
        loss = F.cross_entropy(model(xb), yb)
        loss.backward()
        opt.step()
        loss_value[0] = float(loss)

    while state.step < a.steps and not guard.detached:
        if guard.run(one_step):  # False: a peer died during this step; the survivors are in sync again, repeat it
            state.step += 1
            if state.step % 10 == 0 and kf.current_rank() == 0:
                print(f"step {state.step} np={kf.current_cluster_size()} loss {loss_value[0]:.4f}", flush=True)
    if kf.current_rank() == 0:
        for old, new, secs in guard.failures:
            print(f"lost a worker: {old} -> {new} workers, training again after {secs * 1e3:.0f} ms", flush=True)
        print(f"done: {state.step} steps, final cluster size {kf.current_cluster_size()}", flush=True)


if __name__ == "__main__":
    main()
