"""Synthetic training-throughput benchmark for every optimizer.

Reference: benchmarks/system/benchmark_kungfu_torch.py:58-135 and benchmark_kungfu.py (`--kf-optimizer sync-sgd | sma |
async-sgd`, 10 warm-up batches, `num_iters` x `num_batches_per_iter`, "Img/sec per GPU: mean +-1.96 sigma" and a
`RESULT:` line). Differences: bf16 + channels_last by default, device-timed (CUDA events), max over ranks.

    torchrun --nproc-per-node 8 benchmarks/system/benchmark_kungfu_torch.py --kf-optimizer sma --model resnet50
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200 import models  # noqa: E402
from kungfu_b200.models import resnet  # noqa: E402
from kungfu_b200.runtime import get_runtime  # noqa: E402
from kungfu_b200.utils.timing import max_over_ranks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--batch-size", type=int, default=256)
    ap.add_argument("--kf-optimizer", default="sync-sgd", choices=["sync-sgd", "sma", "async-sgd", "noise-scale", "variance"])
    ap.add_argument("--num-warmup-batches", type=int, default=10)
    ap.add_argument("--num-batches-per-iter", type=int, default=10)
    ap.add_argument("--num-iters", type=int, default=5)
    ap.add_argument("--momentum", type=float, default=0.9)
    ap.add_argument("--fp32", action="store_true")
    a = ap.parse_args()
    rt = get_runtime()
    dev = rt.cuda_index()
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1)
    n = rt.size
    dtype = torch.float32 if a.fp32 else torch.bfloat16
    model = resnet.to_b200(models.build(a.model), dtype, "cuda")  # resnet18/50/101, vgg11/16/19, inception_v3
    side = models.input_size(a.model)
    base = torch.optim.SGD(model.parameters(), lr=0.01 * n, momentum=a.momentum)
    named = list(model.named_parameters())
    opt = {
        "sync-sgd": lambda: kf.optimizers.SynchronousSGDOptimizer(base, named),
        "sma": lambda: kf.optimizers.SynchronousAveragingOptimizer(base, named),
        "async-sgd": lambda: kf.optimizers.PairAveragingOptimizer(base, named),
        "noise-scale": lambda: kf.optimizers.MonitorGradientNoiseScaleOptimizer(base, named, device_batch_size=a.batch_size),
        "variance": lambda: kf.optimizers.MonitorGradientVarianceOptimizer(base, named),
    }[a.kf_optimizer]()
    kf.broadcast_parameters(model.state_dict())
    x = torch.randn(a.batch_size, 3, side, side, device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (a.batch_size,), device="cuda")

    def step():
        opt.zero_grad()
        F.cross_entropy(model(x).float(), y).backward()
        opt.step()

    for _ in range(a.num_warmup_batches):
        step()
    rates = []
    for _ in range(a.num_iters):
        torch.cuda.synchronize()
        rt.control.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(a.num_batches_per_iter):
            step()
        t1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(t0.elapsed_time(t1))
        rates.append(a.batch_size * a.num_batches_per_iter / (ms / 1e3))
    if opt.engine is not None:
        opt.engine.check()
    if rt.rank == 0:
        r = np.array(rates)
        extra = {}
        if a.kf_optimizer == "noise-scale":
            extra["noise_scale"] = opt.noise_scale()
        if a.kf_optimizer == "variance":
            extra["variance"] = opt.variance()
        print("Img/sec per GPU: %.1f +-%.1f" % (r.mean(), 1.96 * r.std()))
        print("Total img/sec on %d GPU(s): %.1f +-%.1f" % (n, n * r.mean(), n * 1.96 * r.std()))
        print("RESULT: %.1f +-%.1f %s" % (n * r.mean(), n * 1.96 * r.std(), json.dumps(
            {"np": n, "model": a.model, "kf_optimizer": a.kf_optimizer, "batch_size": a.batch_size, "dtype": str(dtype), **extra})))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
