"""Elastic training with a step schedule (reference: examples/torch_elastic/torch_mnist_example.py:66-82 — a dict
{step: new_size} + `kfpy.resize`, then all-reduce-max of the step and broadcast of model + optimizer state).

    kungfu-run -w -np 1 -H 127.0.0.1:4 -builtin-config-port 9100 python examples/torch_elastic/torch_mnist_elastic.py
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "../..")))  # run from a checkout without installing

import torch
import torch.nn.functional as F

import kungfu_b200.python as kfp
import kungfu_b200.torch as kf
from kungfu_b200.elastic import ElasticState, sync_after_resize
from kungfu_b200.models.slp import SLP, synthetic_mnist

SCHEDULE = {10: 2, 20: 4, 40: 2, 50: 1}
MAX_STEP = 60


def main():
    model = SLP()
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), named_parameters=model.named_parameters())
    state = sync_after_resize(ElasticState(), model, opt)  # late joiners adopt the survivors' progress and weights
    x, y = synthetic_mnist(2000)
    while state.step < MAX_STEP:
        n, r = kf.current_cluster_size(), kf.current_rank()
        idx = (torch.arange(32) + (state.step * n + r) * 32) % len(x)
        opt.zero_grad()
        F.cross_entropy(model(x[idx]), y[idx]).backward()
        opt.step()
        state.step += 1
        if state.step in SCHEDULE:
            changed, detached = kfp.resize(SCHEDULE[state.step])
            if detached:
                print(f"rank {r} detached at step {state.step}")
                return
            if changed:
                state = sync_after_resize(state, model, opt)
                print(f"step {state.step}: now rank {kf.current_rank()} of {kf.current_cluster_size()}")
    if kf.current_rank() == 0:
        print("done, accuracy", (model(x).argmax(1) == y).float().mean().item())


if __name__ == "__main__":
    main()
