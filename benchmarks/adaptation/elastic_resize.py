"""Elastic resize 4 -> 8 -> 4 GPUs with the gradient-noise-scale monitor (BASELINE.json config #5): img/s per phase and
the latency of each `resize_cluster` (survivors keep state, newcomers receive it by broadcast, every symmetric arena
re-discovers its peers and rebuilds mappings + multicast).

Reference: benchmarks/system/benchmark_kungfu_elastic.py:160-196, benchmarks/adaptation/bench-adaptation.sh.

    kungfu-run -w -np 4 -H 127.0.0.1:8 -builtin-config-port 9100 python benchmarks/adaptation/elastic_resize.py
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import kungfu_b200.python as kfp  # noqa: E402
import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200.elastic import ElasticState, sync_after_resize  # noqa: E402
from kungfu_b200.models import resnet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--schedule", default="4:10,8:10,4:10")
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--steps", type=int, default=0, help="with --follow-config-server: how many steps to run")
    ap.add_argument("--follow-config-server", action="store_true",
                    help="do not drive the size from --schedule: after every step adopt whatever cluster the config server holds "
                         "(an external client / operator resizes the job; reference: benchmark_kungfu_elastic.py:160-196)")
    a = ap.parse_args()
    from kungfu_b200.ops.adapt import StepBasedSchedule

    sch = StepBasedSchedule(a.schedule)
    use_gpu = torch.cuda.is_available() and kf.get_cuda_index() >= 0
    if use_gpu:
        torch.cuda.set_device(kf.get_cuda_index())
    dev = "cuda" if use_gpu else "cpu"
    dtype = torch.bfloat16 if use_gpu else torch.float32
    torch.manual_seed(0)
    model = getattr(resnet, a.model)().to(dev, dtype)
    if use_gpu:
        model = model.to(memory_format=torch.channels_last)
    opt = kf.optimizers.MonitorGradientNoiseScaleOptimizer(torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9), model.named_parameters(),
                                                           device_batch_size=a.batch_size)
    state = sync_after_resize(ElasticState(), model, opt)
    x = torch.randn(a.batch_size, 3, 64 if not use_gpu else 224, 64 if not use_gpu else 224, device=dev, dtype=dtype)
    if use_gpu:
        x = x.contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (a.batch_size,), device=dev)
    t_phase, n_phase = time.time(), 0
    total_steps = a.steps if a.follow_config_server and a.steps > 0 else sch.total_steps
    while state.step < total_steps:
        opt.zero_grad()
        F.cross_entropy(model(x).float(), y).backward()
        opt.step()
        state.step += 1
        n_phase += 1
        if a.follow_config_server:
            t0 = time.time()
            changed, detached = kfp.resize()  # consensus on the server's current cluster, adopt it if it differs
            if detached:
                return
            if changed:
                state = sync_after_resize(state, model, opt)
                if kf.current_rank() == 0:
                    print(json.dumps({"resize_to": kf.current_cluster_size(), "latency_s": round(time.time() - t0, 3), "step": state.step}), flush=True)
            continue
        want = sch(state.step) if state.step < sch.total_steps else kf.current_cluster_size()
        if want != kf.current_cluster_size():
            if use_gpu:
                torch.cuda.synchronize()
            if kf.current_rank() == 0:
                dt = time.time() - t_phase
                print(json.dumps({"phase_np": kf.current_cluster_size(), "img_per_s": round(n_phase * a.batch_size * kf.current_cluster_size() / dt, 1),
                                  "noise_scale": opt.noise_scale()}), flush=True)
            t0 = time.time()
            changed, detached = kfp.resize(want)
            if detached:
                return
            if changed:
                state = sync_after_resize(state, model, opt)
                if kf.current_rank() == 0:
                    print(json.dumps({"resize_to": kf.current_cluster_size(), "latency_s": round(time.time() - t0, 3)}), flush=True)
            t_phase, n_phase = time.time(), 0
    if kf.current_rank() == 0:
        print(json.dumps({"done": True, "steps": state.step}), flush=True)


if __name__ == "__main__":
    main()
