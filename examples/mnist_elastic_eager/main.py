"""Elastic MNIST written against the LOW-LEVEL API only: the training loop itself keeps the job consistent across resizes.

Reference: examples/mnist_elastic_eager/main.py:88-127 (TF2 eager loop: a {step: new_size} schedule, `resize(new_size)`, then
`sync_offsets` + `sync_model` by broadcast when the membership changed, exit when `detached()`, stop after a sample budget).
This is the torch form of the same loop; what the helpers of `kungfu_b200.elastic` do (`ElasticState`, `sync_after_resize`) is spelled
out here with three primitives: `kfp.resize`, `kfp.all_reduce_int_max`, `kf.broadcast_parameters`.

    kungfu-run -w -np 1 -H 127.0.0.1:4 -builtin-config-port 9100 python examples/mnist_elastic_eager/main.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "../..")))  # run from a checkout without installing

import torch
import torch.nn.functional as F

import kungfu_b200.python as kfp
import kungfu_b200.torch as kf
from kungfu_b200.models.slp import synthetic_mnist


def parse_args():
    p = argparse.ArgumentParser(description="elastic MNIST, eager loop")
    p.add_argument("--batch-size", type=int, default=50)
    p.add_argument("--samples", type=int, default=6000, help="size of the (synthetic, MNIST-shaped) training set")
    p.add_argument("--num-epochs", type=float, default=1.0)
    p.add_argument("--learning-rate", type=float, default=0.05)
    p.add_argument("--schedule", default="20:2,40:3,60:4,80:2,100:3,110:1", help="step:new_size,...")
    return p.parse_args()


def build_model():
    # the reference's small CNN (two 3x3 convolutions, max-pool, dropout, two dense layers)
    return torch.nn.Sequential(torch.nn.Conv2d(1, 8, 3), torch.nn.ReLU(), torch.nn.Conv2d(8, 16, 3), torch.nn.ReLU(),
                               torch.nn.MaxPool2d(2), torch.nn.Dropout(0.25), torch.nn.Flatten(), torch.nn.Linear(16 * 12 * 12, 64), torch.nn.ReLU(),
                               torch.nn.Dropout(0.5), torch.nn.Linear(64, 10))


def sync_offsets(global_step, trained_samples):
    """Everybody adopts the furthest progress: survivors agree anyway, a newcomer starts from zero (the reference broadcasts from rank 0
    and notes `TODO: use all_reduce with max op` — this is that)."""
    return kfp.all_reduce_int_max(global_step), kfp.all_reduce_int_max(trained_samples)


def sync_model(model, opt):
    kf.broadcast_parameters(model.state_dict())
    state = opt.state_dict()["state"]
    for k in sorted(state):
        kf.broadcast_parameters({f"opt.{k}.{n}": v for n, v in sorted(state[k].items()) if torch.is_tensor(v)})


def train(args):
    schedule = {int(a): int(b) for a, b in (kv.split(":") for kv in args.schedule.split(",") if kv)}
    x, y = synthetic_mnist(args.samples)
    x = x.reshape(len(x), 1, 28, 28)
    torch.manual_seed(1234)  # identical initial weights everywhere; late joiners get the current ones by broadcast
    model = build_model()
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(model.parameters(), lr=args.learning_rate), named_parameters=model.named_parameters())
    total_samples = int(args.samples * args.num_epochs)
    global_step, trained_samples, need_sync, resizes = 0, 0, True, []
    while True:
        if need_sync:  # first step of every membership: agree on the progress, then on the weights
            global_step, trained_samples = sync_offsets(global_step, trained_samples)
            sync_model(model, opt)
            need_sync = False
        if trained_samples >= total_samples:
            break
        n, r = kf.current_cluster_size(), kf.current_rank()
        idx = (torch.arange(args.batch_size) + trained_samples + r * args.batch_size) % len(x)  # this worker's shard of the global batch
        opt.zero_grad()
        loss = F.cross_entropy(model(x[idx]), y[idx])
        loss.backward()
        opt.step()
        global_step += 1
        trained_samples += n * args.batch_size
        if r == 0 and global_step % 10 == 0:
            print(f"step: {global_step} np: {n} loss: {loss.item():.4f}", flush=True)
        if global_step in schedule:
            changed, detached = kfp.resize(schedule[global_step])
            if detached:
                print(f"detached after step {global_step}", flush=True)
                return
            if changed:
                resizes.append((global_step, n, kf.current_cluster_size()))
                need_sync = True
    if kf.current_rank() == 0:
        model.eval()
        acc = (model(x).argmax(1) == y).float().mean().item()
        print(f"done: steps={global_step} trained={trained_samples} np={kf.current_cluster_size()} resizes={resizes} accuracy={acc:.4f}", flush=True)


if __name__ == "__main__":
    print("main started", flush=True)
    train(parse_args())
    print("main finished", flush=True)
