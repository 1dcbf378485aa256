"""Synthetic training-throughput benchmark of a job whose size is changed from OUTSIDE while it runs.

Reference: benchmarks/system/benchmark_kungfu_elastic.py:160-196 — after every iteration the job asks the config server for the
cluster (`resize_cluster_from_url`); after a change the step counter is agreed on (all-reduce max), the variables are broadcast, and
the per-iteration "img/sec" lines continue; detached workers leave. The resizes come from an operator or from a resizing client
(`python -m kungfu_b200.cmd.config_client`, the counterpart of kungfu-config-client-example), or from `--schedule` for a self-contained run.

    kungfu-run -w -np 2 -H 127.0.0.1:8 -builtin-config-port 9100 python benchmarks/system/benchmark_kungfu_elastic.py --num-iters 60 &
    python -m kungfu_b200.cmd.config_client --server http://127.0.0.1:9100 --max-size 8 --period 2

Prints, per membership, the mean rate and the cost of the change (resize + state re-sync), and a final `RESULT:` line like the other
system benchmarks. Runs on the GPU plane when there is a GPU (bf16, channels_last) and on the host plane otherwise (`--no-cuda`, small images).
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

import kungfu_b200.python as kfp  # noqa: E402
import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200 import models  # noqa: E402
from kungfu_b200.elastic import ElasticState, sync_after_resize  # noqa: E402


def build_optimizer(name, base, named, batch_size):
    if name == "sync-sgd":
        return kf.optimizers.SynchronousSGDOptimizer(base, named)
    if name == "sma":
        return kf.optimizers.SynchronousAveragingOptimizer(base, named)
    if name == "async-sgd":
        return kf.optimizers.PairAveragingOptimizer(base, named)
    if name == "noise-scale":
        return kf.optimizers.MonitorGradientNoiseScaleOptimizer(base, named, device_batch_size=batch_size)
    raise SystemExit(f"unknown --kf-optimizer {name}")


def main():
    ap = argparse.ArgumentParser(description="elastic synthetic benchmark", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--model", default="resnet50", help="resnet18/50/101, vgg16, inception_v3, or slp")
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--num-warmup-batches", type=int, default=3)
    ap.add_argument("--num-batches-per-iter", type=int, default=1)
    ap.add_argument("--num-iters", type=int, default=30)
    ap.add_argument("--no-cuda", action="store_true")
    ap.add_argument("--image-size", type=int, default=0, help="0: the model's ImageNet crop on a GPU, 64 on the host plane")
    ap.add_argument("--kf-optimizer", default="sync-sgd", choices=["sync-sgd", "sma", "async-sgd", "noise-scale"])
    ap.add_argument("--optimizer", default="sgd", choices=["sgd", "adam"])
    ap.add_argument("--schedule", default="", help="iter:size,… — rank 0 proposes these sizes itself instead of waiting for an outside change")
    a = ap.parse_args()

    use_gpu = (not a.no_cuda) and torch.cuda.is_available() and kf.get_cuda_index() >= 0
    if use_gpu:
        torch.cuda.set_device(kf.get_cuda_index())
    dev, dtype = ("cuda", torch.bfloat16) if use_gpu else ("cpu", torch.float32)
    torch.manual_seed(0)
    if a.model == "slp":
        from kungfu_b200.models.slp import SLP

        model = SLP().to(dev, dtype)
        x = torch.randn(a.batch_size, 784, device=dev, dtype=dtype)
        y = torch.randint(0, 10, (a.batch_size,), device=dev)
    else:
        model = models.build(a.model).to(dev, dtype)
        side = a.image_size or (models.input_size(a.model) if use_gpu else 64)
        x = torch.randn(a.batch_size, 3, side, side, device=dev, dtype=dtype)
        if use_gpu:
            model = model.to(memory_format=torch.channels_last)
            x = x.contiguous(memory_format=torch.channels_last)
        y = torch.randint(0, 1000, (a.batch_size,), device=dev)
    base = (torch.optim.SGD(model.parameters(), lr=0.01) if a.optimizer == "sgd" else torch.optim.Adam(model.parameters(), lr=1e-3))
    opt = build_optimizer(a.kf_optimizer, base, list(model.named_parameters()), a.batch_size)
    schedule = {int(k): int(v) for k, v in (it.split(":") for it in a.schedule.split(",") if it)}

    def train_batch():
        opt.zero_grad()
        F.cross_entropy(model(x).float(), y).backward()
        opt.step()

    def sync(dev_sync=True):
        if use_gpu and dev_sync:
            torch.cuda.synchronize()

    state = ElasticState()
    need_sync, warm = True, False
    rates, phases, changes = [], [], []  # phases: [np, first iter, rates…]
    device_name = "GPU" if use_gpu else "CPU"
    while state.step < a.num_iters:
        if need_sync:
            t0 = time.time()
            before = state.step
            state = sync_after_resize(state, model, opt)  # step = max over members, model / optimizer state from rank 0
            sync()
            if state.step != before:
                print(f"sync step : {before} -> {state.step}", flush=True)
            print(f"state sync took {time.time() - t0:.3f}s (np={kf.current_cluster_size()})", flush=True)
            if changes and changes[-1].get("sync_s") is None:
                changes[-1]["sync_s"] = round(time.time() - t0, 4)
            need_sync = False
            phases.append({"np": kf.current_cluster_size(), "from_iter": state.step, "rates": []})
            if state.step >= a.num_iters:
                break
        if not warm:
            # warm-up batches are collective: only the members the job starts with run them, a newcomer (step > 0 after the sync) may not
            if state.step == 0:
                for _ in range(a.num_warmup_batches):
                    train_batch()
                sync()
            warm = True
        state.step += 1
        t0 = time.time()
        for _ in range(a.num_batches_per_iter):
            train_batch()
        sync()
        dt = time.time() - t0
        img_sec = a.batch_size * a.num_batches_per_iter / dt
        n = kf.current_cluster_size()
        print(f"Iter #{state.step}: {img_sec:.1f} img/sec per {device_name} (np={n})", flush=True)
        rates.append(img_sec)
        phases[-1]["rates"].append(img_sec)
        t0 = time.time()
        want = schedule.get(state.step)
        changed, detached = kfp.resize(want) if want is not None else kfp.resize(None)
        if changed:
            changes.append({"iter": state.step, "from": n, "to": None if detached else kf.current_cluster_size(), "resize_s": round(time.time() - t0, 4),
                            "sync_s": None})
            print(f"resize at iter {state.step}: {n} -> {'detached' if detached else kf.current_cluster_size()} took {time.time() - t0:.3f}s", flush=True)
        if detached:
            print(f"rank detached at iter {state.step}", flush=True)
            return
        need_sync = changed
    if kf.current_rank() != 0:
        return
    mean, conf = float(np.mean(rates)) if rates else 0.0, float(1.96 * np.std(rates)) if rates else 0.0
    print(f"Img/sec per {device_name}: {mean:.1f} +-{conf:.1f}")
    for ph in phases:
        if ph["rates"]:
            print(f"  np={ph['np']}: {len(ph['rates'])} iters from #{ph['from_iter'] + 1}, {np.mean(ph['rates']):.1f} img/sec per {device_name}, "
                  f"{np.mean(ph['rates']) * ph['np']:.1f} whole job")
    attrs = {"np": kf.current_cluster_size(), "strategy": os.getenv("KUNGFU_ALLREDUCE_STRATEGY"), "bs": a.batch_size, "model": a.model,
             "kf-opt": a.kf_optimizer, "device": device_name, "changes": changes}
    print("RESULT: %f +-%f %s" % (mean, conf, json.dumps(attrs, separators=(",", ":"))))


if __name__ == "__main__":
    main()
