#!/usr/bin/env python
"""Scaling benchmark with an optional resize schedule: synthetic ImageNet-shaped training whose cluster size follows
`--resize-schedule size:steps,...`, per-step throughput logged by `LogPerfHook`, one RESULT line per phase.

Reference: benchmarks/scaling/benchmark_kungfu_scaling.py (TF `simple | monitored | estimator` loops + `KungFuElasticTrainHook`,
`--resize-schedule 10:2,100:0`, `--show-training-throughput`). Here: one PyTorch loop, S-SGD optimizer, elastic hook.

    kungfu-run -w -np 1 -H 127.0.0.1:8 -builtin-config-port 9100 python benchmarks/scaling/benchmark_kungfu_scaling.py \
        --elastic --resize-schedule 1:10,2:10,4:10,8:10 --train-steps 40 [--model resnet50 --batch-size 64]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import debug_hooks  # noqa: E402

import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200 import models  # noqa: E402
from kungfu_b200.elastic import KungFuElasticTrainHook  # noqa: E402
from kungfu_b200.utils import _log_event  # noqa: E402


def parse_args():
    p = argparse.ArgumentParser(description="Synthetic scaling benchmark", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--model", default="slp", help="slp | resnet18 | resnet50 | vgg16 | inception_v3")
    p.add_argument("--elastic", action="store_true")
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--train-steps", type=int, default=10)
    p.add_argument("--resize-schedule", default="1:10", help="size:steps,size:steps,... as in StepBasedSchedule (used when --elastic)")
    p.add_argument("--model-dir", default="")
    p.add_argument("--show-training-throughput", action="store_true")
    return p.parse_args()


def build(a, dev):
    if a.model == "slp":
        from kungfu_b200.models.slp import SLP

        return SLP(zero_init=False).to(dev), (a.batch_size, 1, 28, 28), 10
    side = models.input_size(a.model)
    dtype = torch.bfloat16 if dev != "cpu" else torch.float32
    m = models.build(a.model).to(device=dev, dtype=dtype, memory_format=torch.channels_last)
    return m, (a.batch_size, 3, side, side), 1000


def main():
    a = parse_args()
    _log_event("BEGIN :: main")
    dev = "cpu"
    if torch.cuda.is_available() and kf.get_cuda_index() >= 0:
        torch.cuda.set_device(kf.get_cuda_index())
        dev = "cuda"
    model, shape, classes = build(a, dev)
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(model.parameters(), lr=0.01), named_parameters=model.named_parameters())
    x = torch.rand(shape, device=dev, dtype=next(model.parameters()).dtype)
    y = torch.randint(0, classes, (a.batch_size,), device=dev)
    hook = KungFuElasticTrainHook(a.resize_schedule, a.train_steps, model, opt, a.model_dir or None) if a.elastic else None
    if hook:
        hook.begin()
    else:
        kf.broadcast_parameters(model.state_dict())
    perf = debug_hooks.LogPerfHook(a.batch_size, warmup_steps=2, quiet=not a.show_training_throughput)
    size = kf.current_cluster_size()
    step = 0
    while True:
        perf.before_run()
        opt.zero_grad()
        F.cross_entropy(model(x).float(), y).backward()
        opt.step()
        if dev == "cuda":
            torch.cuda.synchronize()
        perf.after_run()
        step += 1
        go_on = hook.after_step() if hook else step < a.train_steps
        if kf.current_cluster_size() != size or not go_on:  # a phase ended: report it, start a new one
            if kf.current_rank() == 0 or not go_on:
                perf.end()
            size = kf.current_cluster_size()
            perf = debug_hooks.LogPerfHook(a.batch_size, warmup_steps=1, quiet=not a.show_training_throughput)
        if not go_on:
            break
    _log_event("END :: main")


if __name__ == "__main__":
    main()
