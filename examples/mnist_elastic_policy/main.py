"""Elastic SLP / MNIST training driven by a policy (reference: examples/mnist_elastic_policy/main.py + run.sh: Estimator + `PolicyHook`
with a `ScheduledElasticPolicy`; here a plain PyTorch loop around the same hook).

    kungfu-run -w -np 1 -H 127.0.0.1:4 -builtin-config-port 9100 python examples/mnist_elastic_policy/main.py --schedule 10:2,20:4,40:1
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "../..")))  # run from a checkout without installing
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch
import torch.nn.functional as F

import kungfu_b200.python as kfp
import kungfu_b200.torch as kf
from kungfu_b200.elastic import PolicyHook
from kungfu_b200.models.slp import SLP, synthetic_mnist

from elastic_policy import ScheduledElasticPolicy


def parse_schedule(s):
    return {int(k): int(v) for k, v in (item.split(":") for item in s.split(",") if item)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--schedule", default="10:2,20:4,40:2,50:1", help="step:size,…")
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--samples", type=int, default=2000, help="size of one epoch")
    ap.add_argument("--num-epochs", type=int, default=3)
    ap.add_argument("--learning-rate", type=float, default=0.1)
    args = ap.parse_args()

    model = SLP()
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(model.parameters(), lr=args.learning_rate), named_parameters=model.named_parameters())
    policy = ScheduledElasticPolicy(parse_schedule(args.schedule), model, opt)
    hook = PolicyHook([policy], epoch_size=args.samples, epoch_num=args.num_epochs, init_batch_size=args.batch_size)
    x, y = synthetic_mnist(args.samples)

    hook.begin()
    while hook.before_step():
        n, r, bs = kf.current_cluster_size(), kf.current_rank(), hook.batch_size
        # every worker takes its own slice of the global batch that starts at the number of samples trained so far
        start = hook.vars["kungfu_trained_samples"] + r * bs
        idx = (torch.arange(bs) + start) % len(x)
        opt.zero_grad()
        loss = F.cross_entropy(model(x[idx]), y[idx])
        loss.backward()
        opt.step()
        hook.after_step()
        if hook.vars["step"] % 10 == 0 and r == 0:
            print(f"step {hook.vars['step']}: np={n} trained={hook.vars['kungfu_trained_samples']} loss={loss.item():.4f}", flush=True)
    hook.end()
    if kfp.detached():
        print(f"detached after step {hook.vars['step']}", flush=True)
    elif kf.current_rank() == 0:
        acc = (model(x).argmax(1) == y).float().mean().item()
        print(f"done: steps={hook.vars['step']} trained={hook.vars['kungfu_trained_samples']} resizes={policy.resizes} accuracy={acc:.4f}", flush=True)


if __name__ == "__main__":
    main()
