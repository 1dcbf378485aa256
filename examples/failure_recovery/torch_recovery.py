"""Failure detection + auto recovery (reference: examples/Failure_recovery_examples/{eager,tf1_mnist_session}.py; design in
docs/monitor_proposal.md). Heartbeats around every batch; a checkpoint per epoch; on `--restart 1` resume from it.

    kungfu-run -np 4 -auto-recover 10s python examples/failure_recovery/torch_recovery.py --n-epochs 5
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "../..")))  # run from a checkout without installing

import argparse
import os

import torch
import torch.nn.functional as F

import kungfu_b200.torch as kf
from kungfu_b200 import checkpoint
from kungfu_b200.cmd import monitor_batch_begin, monitor_batch_end, monitor_epoch_end, monitor_train_end
from kungfu_b200.models.slp import SLP, synthetic_mnist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-epochs", type=int, default=5)
    ap.add_argument("--restart", type=int, default=0)
    ap.add_argument("--ckpt", default="saved_models/torch_recovery/model.pt")
    a = ap.parse_args()
    model = SLP()
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), named_parameters=model.named_parameters())
    meta = checkpoint.load(a.ckpt, model, None) if a.restart else {"epoch": 0}
    x, y = synthetic_mnist(3000)
    n, r = kf.current_cluster_size(), kf.current_rank()
    for e in range(a.n_epochs):
        for s in range(0, 3000 - 60 * n + 1, 60 * n):
            monitor_batch_begin()
            xb, yb = x[s + r * 60:s + (r + 1) * 60], y[s + r * 60:s + (r + 1) * 60]
            opt.zero_grad()
            F.cross_entropy(model(xb), yb).backward()
            opt.step()
            monitor_batch_end()
        monitor_epoch_end()
        checkpoint.save(a.ckpt, model, None, epoch=meta["epoch"] + e + 1)
        if r == 0:
            print(f"epoch {meta['epoch'] + e}: acc {(model(x).argmax(1) == y).float().mean().item():.4f}")
    monitor_train_end()


if __name__ == "__main__":
    main()
