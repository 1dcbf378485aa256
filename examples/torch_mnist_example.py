"""SLP on (synthetic) MNIST with SynchronousSGDOptimizer (reference: examples/torch_mnist_example.py:162-167).

    kungfu-run -np 2 python examples/torch_mnist_example.py --epochs 2
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))  # run from a checkout without installing

import argparse

import torch
import torch.nn.functional as F

import kungfu_b200.torch as kf
from kungfu_b200.models.slp import SLP, synthetic_mnist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--batch-size", type=int, default=50, help="global batch size")
    ap.add_argument("--lr", type=float, default=0.1)
    ap.add_argument("--optimizer", default="sync-sgd", choices=["sync-sgd", "sma", "pair"])
    a = ap.parse_args()
    n, r = kf.current_cluster_size(), kf.current_rank()
    dev = "cpu"
    if torch.cuda.is_available() and kf.get_cuda_index() >= 0:
        torch.cuda.set_device(kf.get_cuda_index())
        dev = "cuda"
    model = SLP().to(dev)
    base = torch.optim.SGD(model.parameters(), lr=a.lr)
    mk = {"sync-sgd": kf.optimizers.SynchronousSGDOptimizer, "sma": kf.optimizers.SynchronousAveragingOptimizer,
          "pair": kf.optimizers.PairAveragingOptimizer}[a.optimizer]
    opt = mk(base, named_parameters=model.named_parameters())
    kf.broadcast_parameters(model.state_dict())
    x, y = synthetic_mnist(6000)
    x, y = x.to(dev), y.to(dev)
    lb = a.batch_size // n
    for epoch in range(a.epochs):
        for s in range(0, len(x) - a.batch_size + 1, a.batch_size):
            xb, yb = x[s + r * lb:s + (r + 1) * lb], y[s + r * lb:s + (r + 1) * lb]
            opt.zero_grad()
            loss = F.cross_entropy(model(xb), yb)
            loss.backward()
            opt.step()
        acc = (model(x).argmax(1) == y).float().mean().item()
        if r == 0:
            print(f"epoch {epoch}: loss {loss.item():.4f} acc {acc:.4f}")


if __name__ == "__main__":
    main()
