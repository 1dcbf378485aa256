#!/usr/bin/env python
"""What online monitoring costs: training throughput with the gradient-variance / gradient-noise-scale monitors against plain S-SGD,
in ONE job (same model, same data, same warm caches), for several monitoring intervals.

    kungfu-run -np 4 python benchmarks/monitoring/benchmark.py --kf-optimizer noise-scale --model resnet50 --batch-size 64
    kungfu-run -np 2 python benchmarks/monitoring/benchmark.py --kf-optimizer both --intervals 1,10 --no-cuda --model slp

Every arm trains `--num-iters` x `--num-batches-per-iter` batches after `--num-warmup-batches`; rank 0 prints one `RESULT:` line per arm
(img/sec over the whole job, 95 % interval over the iterations, overhead against the plain arm, the last monitored value) and one JSON
summary. On the GPU path the monitors are side reductions of kernels that run anyway (the gradient hand-over copy computes the local
sum of squares, the all-reduce epilogue the global one), which is why the expected overhead is ~0; on the CPU path they are two extra small
all-reduces every `interval` steps.

Reference: benchmarks/monitoring/benchmark.py:18-148 (`--kf-optimizer variance|noise-scale`, `--interval`, the horovod-style
"Img/sec per GPU" report) and README.md:1-17. The reference runs one optimizer per job and leaves the comparison to the reader.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200.runtime import get_runtime  # noqa: E402


def build(args, device, dtype):
    if args.model == "slp":  # tiny, for CPU runs
        from kungfu_b200.models.slp import SLP

        model = SLP().to(device)
        x = torch.randn(args.batch_size, 1, 28, 28, device=device)
        y = torch.randint(0, 10, (args.batch_size,), device=device)
        return model, x, y
    from kungfu_b200 import models
    from kungfu_b200.models import resnet

    side = models.input_size(args.model)
    if device.type == "cuda":
        model = resnet.to_b200(models.build(args.model), dtype, "cuda")
        x = torch.randn(args.batch_size, 3, side, side, device=device, dtype=dtype).contiguous(memory_format=torch.channels_last)
    else:
        model = models.build(args.model)
        x = torch.randn(args.batch_size, 3, side, side)
    y = torch.randint(0, 1000, (args.batch_size,), device=device)
    return model, x, y


def make_optimizer(kind, interval, model, args, n):
    base = torch.optim.SGD(model.parameters(), lr=0.01 * n, momentum=args.momentum)
    named = list(model.named_parameters())
    if kind == "sync-sgd":
        return kf.optimizers.SynchronousSGDOptimizer(base, named)
    if kind == "noise-scale":
        return kf.optimizers.MonitorGradientNoiseScaleOptimizer(base, named, device_batch_size=args.batch_size, monitor_interval=interval)
    if kind == "variance":
        return kf.optimizers.MonitorGradientVarianceOptimizer(base, named, monitor_interval=interval)
    raise ValueError(kind)


def run_arm(kind, interval, args, rt, device, dtype):
    """Fresh model + optimizer per arm (identical seed), warm-up, then timed iterations. Returns (per-iteration img/s of the job, value)."""
    torch.manual_seed(1000 + rt.rank)  # every worker sees its own data (identical batches would make every monitor read zero); the
    model, x, y = build(args, device, dtype)  # models are made equal by the broadcast below
    opt = make_optimizer(kind, interval, model, args, rt.size)
    kf.broadcast_parameters(model.state_dict())
    cuda = device.type == "cuda"

    def step():
        opt.zero_grad()
        out = model(x)
        F.cross_entropy(out.float() if out.dtype != torch.float32 else out, y).backward()
        opt.step()

    for _ in range(args.num_warmup_batches):
        step()
    rates = []
    for _ in range(args.num_iters):
        if cuda:
            torch.cuda.synchronize()
        rt.control.barrier()
        t0 = time.perf_counter()
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for _ in range(args.num_batches_per_iter):
            step()
        if cuda:
            e1.record()
            torch.cuda.synchronize()
            secs = e0.elapsed_time(e1) / 1e3
        else:
            secs = time.perf_counter() - t0
        from kungfu_b200.utils.timing import max_over_ranks

        secs = max_over_ranks(secs)
        rates.append(rt.size * args.batch_size * args.num_batches_per_iter / secs)
    value = None
    if kind == "noise-scale":
        value = float(opt.noise_scale())
    elif kind == "variance":
        value = float(opt.variance())
    eng = getattr(opt, "engine", None)
    if eng is not None:
        eng.check()
    del opt, model
    return np.array(rates), value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--kf-optimizer", default="both", choices=["variance", "noise-scale", "both"])
    ap.add_argument("--interval", type=int, default=None, help="one monitoring interval (reference flag)")
    ap.add_argument("--intervals", default="1,10", help="comma-separated monitoring intervals to sweep (ignored when --interval is given)")
    ap.add_argument("--num-warmup-batches", type=int, default=10)
    ap.add_argument("--num-batches-per-iter", type=int, default=10)
    ap.add_argument("--num-iters", type=int, default=5)
    ap.add_argument("--momentum", type=float, default=0.9)
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--no-cuda", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rt = get_runtime()
    use_cuda = not args.no_cuda and torch.cuda.is_available() and rt.has_gpu()
    if use_cuda:
        torch.cuda.set_device(rt.cuda_index())
        torch.backends.cudnn.benchmark = True
    device = torch.device("cuda", rt.cuda_index()) if use_cuda else torch.device("cpu")
    dtype = torch.float32 if (args.fp32 or not use_cuda) else torch.bfloat16
    intervals = [args.interval] if args.interval else [int(v) for v in args.intervals.split(",") if v]
    kinds = ["variance", "noise-scale"] if args.kf_optimizer == "both" else [args.kf_optimizer]

    def log(s):
        if rt.rank == 0:
            print(s, flush=True)

    log(f"Model: {args.model}  batch size {args.batch_size} per worker  workers {rt.size}  device {device.type}  dtype {dtype}")
    base, _ = run_arm("sync-sgd", 0, args, rt, device, dtype)
    rows = [{"arm": "sync-sgd", "interval": None, "img_per_sec": round(float(base.mean()), 1), "ci95": round(float(1.96 * base.std()), 1),
             "overhead_pct": 0.0, "value": None}]
    log("RESULT: sync-sgd %.1f +-%.1f img/sec on %d worker(s) (no monitor)" % (base.mean(), 1.96 * base.std(), rt.size))
    for kind in kinds:
        for iv in intervals:
            r, value = run_arm(kind, iv, args, rt, device, dtype)
            over = 100.0 * (float(base.mean()) / float(r.mean()) - 1.0)
            rows.append({"arm": kind, "interval": iv, "img_per_sec": round(float(r.mean()), 1), "ci95": round(float(1.96 * r.std()), 1),
                         "overhead_pct": round(over, 2), "value": value})
            log("RESULT: %s interval=%d %.1f +-%.1f img/sec on %d worker(s), overhead %.2f %%, last %s = %s"
                % (kind, iv, r.mean(), 1.96 * r.std(), rt.size, over, kind.replace("-", " "), "%.6g" % value if value is not None else "n/a"))
    if rt.rank == 0:
        summary = {"bench": "monitoring", "model": args.model, "batch_size": args.batch_size, "np": rt.size, "device": device.type,
                   "dtype": str(dtype), "rows": rows}
        print(json.dumps(summary), flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(summary, f, indent=1)
    if use_cuda:
        torch.cuda.synchronize()
    rt.control.barrier()


if __name__ == "__main__":
    main()
