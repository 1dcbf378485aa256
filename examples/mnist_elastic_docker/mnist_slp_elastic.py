"""Elastic SLP training whose cluster is managed from OUTSIDE: the program only follows the config server.

Reference: examples/mnist_elastic_docker/mnist_slp_estimator.py (an Estimator with `KungFuElasticTrainHook`; a cluster manager starts /
reclaims nodes and writes the membership into the config server with its config client). Torch form: every step ends with
`resize()` with no argument — "adopt whatever the config server holds now" — and the sample budget decides when training is over, so the
job finishes no matter which sequence of sizes the manager chose.

    # one machine, virtual nodes (no docker): config server + one launcher per node + a random resizer
    python experiments/cluster_manager.py --nodes 3 --slots 2 --duration 14 --period 3.5 --late-node-after 4 -- \
        python examples/mnist_elastic_docker/mnist_slp_elastic.py --step-ms 30 --num-epochs 100
    # docker compose on a private subnet
    python benchmarks/adaptation/gen_compose.py --nodes 4 --slots 2 > docker-compose.yaml && docker compose up
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "../..")))  # run from a checkout without installing

import torch
import torch.nn.functional as F

import kungfu_b200.python as kfp
import kungfu_b200.torch as kf
from kungfu_b200.elastic import ElasticState, sync_after_resize
from kungfu_b200.models.slp import SLP, synthetic_mnist


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--batch-size", type=int, default=50)
    p.add_argument("--samples", type=int, default=6000)
    p.add_argument("--num-epochs", type=float, default=2.0)
    p.add_argument("--learning-rate", type=float, default=0.1)
    p.add_argument("--step-ms", type=float, default=0.0, help="artificial time per step (lets an outside manager get a word in on a toy model)")
    a = p.parse_args()
    import time

    model = SLP()
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(model.parameters(), lr=a.learning_rate), named_parameters=model.named_parameters())
    state = sync_after_resize(ElasticState(), model, opt)  # a node started late adopts the survivors' progress and weights
    x, y = synthetic_mnist(a.samples)
    total = int(a.samples * a.num_epochs)
    sizes = []
    while state.trained_samples < total:
        n, r = kf.current_cluster_size(), kf.current_rank()
        idx = (torch.arange(a.batch_size) + state.trained_samples + r * a.batch_size) % len(x)
        opt.zero_grad()
        F.cross_entropy(model(x[idx]), y[idx]).backward()
        opt.step()
        state.step += 1
        state.trained_samples += n * a.batch_size
        if not sizes or sizes[-1] != n:
            sizes.append(n)
        if a.step_ms > 0:
            time.sleep(a.step_ms / 1e3)
        changed, detached = kfp.resize(None)  # follow the config server
        if detached:
            print(f"detached after step {state.step}", flush=True)
            return
        if changed:
            state = sync_after_resize(state, model, opt)
    if kf.current_rank() == 0:
        acc = (model(x).argmax(1) == y).float().mean().item()
        print(f"done: steps={state.step} trained={state.trained_samples} sizes_seen_by_rank0={sizes} accuracy={acc:.4f}", flush=True)


if __name__ == "__main__":
    main()
