#!/usr/bin/env python
"""Headline benchmark: ResNet-50 SynchronousSGDOptimizer training throughput (images/sec, whole job).

Mirrors the reference's synthetic benchmark (benchmarks/system/benchmark_kungfu_torch.py:58-135: resnet50, synthetic
ImageNet-shaped batch, SGD lr=0.01*np wrapped in SynchronousSGDOptimizer, warm-up then timed batches) on the
BASELINE.json config "ResNet-50 SynchronousSGDOptimizer bf16", per-GPU batch 256 (the batch of the published headline
figure), but device-timed with CUDA events, max over ranks.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      bench.py --gpus N --steps K --warmup W

--impl ours       our framework (fused sm_100a reduce-scatter + optimizer + all-gather kernels)     [default]
--impl nccl       the same-box baseline: SAME model, fused BN kernels, bf16 + fp32-master optimizer kernel, bucket plan, comm
                  stream and CUDA graph — only the collective is `ncclAllReduce` (torch.distributed) instead of our kernels
                  ("a path that only calls NCCL is the baseline, not the product")
--impl reference  the unmodified reference from baseline/_ref (cannot be installed offline: needs Go + TensorFlow)
--optimizer ssgd | sma    BASELINE.json configs #2 (SynchronousSGDOptimizer, default) and #3 (SynchronousAveragingOptimizer)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

BASELINE_IMG_S = 5507.0  # ResNet-50 S-SGD, 16 x V100, total batch 4096 (256/GPU), whole cluster


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "nccl", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch size")
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--momentum", type=float, default=0.0)
    ap.add_argument("--optimizer", default="ssgd", choices=["ssgd", "sma"])
    ap.add_argument("--trace", default="", help="after the timed runs: CUPTI kernel timeline of 3 steps per rank -> <prefix>.rank<r>.json")
    return ap.parse_args()


def reference_arm(args) -> int:
    """The reference (kungfu 0.2.2) builds a Go c-archive and TensorFlow op kernels at install time; neither a Go
    toolchain nor TensorFlow exists offline here, so `pip install /root/reference` fails."""
    ref = os.path.join(os.path.dirname(os.path.abspath(__file__)), "baseline", "_ref")
    why = "reference needs Go>=1.11 and TensorFlow at build time; neither is installable offline (pip build fails in setup.py cmake_tf_ext_flags: No module named 'tensorflow')"
    if os.path.isdir(os.path.join(ref, "kungfu")):
        sys.path.insert(0, ref)
        try:
            import kungfu.torch  # noqa: F401

            why = "reference package present but its native libkungfu is missing"
        except Exception as e:  # pragma: no cover
            why = f"reference import failed: {type(e).__name__}: {e}"
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"impl": "reference", "unavailable": why}))
    return 0


def _trace(args, trainer, rt) -> None:
    """Real (not replayed) kernel timeline of 3 steps on every rank: start / duration / stream of each kernel, so that the
    collective kernels can be placed against backward and the exposed tail of a step can be read off."""
    import torch
    from torch.profiler import ProfilerActivity, profile

    torch.cuda.synchronize()
    rt.control.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            trainer.device_step()
        torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if getattr(e, "device_type", None) is not None and e.device_type.name == "CUDA" and e.device_time_total > 0:
            tr = e.time_range
            evs.append({"name": e.name[:160], "ts": tr.start, "dur": tr.end - tr.start})
    evs.sort(key=lambda x: x["ts"])
    path = f"{args.trace}.rank{rt.rank}.json"
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        json.dump({"n_gpus": rt.size, "impl": args.impl, "optimizer": args.optimizer, "events": evs}, f)
    try:
        prof.export_chrome_trace(f"{args.trace}.rank{rt.rank}.chrome.json")
    except Exception:
        pass
    rt.control.barrier()


def main() -> int:
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)

    import torch

    import kungfu_b200.torch as kf
    from kungfu_b200.models import resnet
    from kungfu_b200.runtime import get_runtime
    from kungfu_b200.trainer import ImageTrainer
    from kungfu_b200.utils.timing import ClockSampler, max_over_ranks

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device", "impl": args.impl}))
        return 1
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # self-launch under torchrun so that `python bench.py --gpus N` also works
            import subprocess

            port = 29000 + (os.getpid() % 2000)
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd)
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    rt = get_runtime()
    dev_index = rt.cuda_index()
    torch.cuda.set_device(dev_index)
    device = torch.device("cuda", dev_index)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1234)  # identical initial weights on every rank (stands in for broadcast_parameters)
    N = rt.size
    B = args.batch

    model = getattr(resnet, args.model)()
    model = resnet.to_b200(model, torch.bfloat16, device)
    lr = 0.01 * N
    base_opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=args.momentum)

    backend = None
    if args.impl == "nccl":
        backend = "nccl"
        if N > 1:
            import torch.distributed as dist

            # CUDA-graph capture of NCCL collectives: the watchdog's event queries must not run against a capturing stream
            os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
            os.environ.setdefault("NCCL_ASYNC_ERROR_HANDLING", "0")
            if not dist.is_initialized():
                dist.init_process_group("nccl", device_id=device)
            warm = torch.ones(1 << 20, device=device)
            dist.all_reduce(warm)  # communicator + channels exist before anything is captured
            torch.cuda.synchronize()
    if args.optimizer == "ssgd":
        opt = kf.optimizers.SynchronousSGDOptimizer(base_opt, named_parameters=model.named_parameters(), backend=backend)
    else:
        opt = kf.optimizers.SynchronousAveragingOptimizer(base_opt, named_parameters=model.named_parameters(), backend=backend)
    kf.broadcast_parameters(model.state_dict())
    train_model = model
    if os.environ.get("KUNGFU_BENCH_NO_COMM") == "1" and N > 1:
        # diagnostic arm: N processes with their arenas connected and peer-mapped, but every rank trains on its own (the engine is
        # told it is alone) — separates "what the collectives cost" from "what having peer mappings / N busy GPUs costs"
        opt.engine.nranks = 1
        opt.engine._native = None

    trainer = ImageTrainer(train_model, opt, B, 224, torch.bfloat16, use_graph=not args.no_graph, warmup_iters=3 if backend is None else 12,
                           device=device, capture_error_mode="thread_local" if backend == "nccl" else "global")
    # synthetic ImageNet-shaped data
    g = torch.Generator().manual_seed(rt.rank)
    host_images = [torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(2)]
    host_labels = [torch.randint(0, 1000, (B,), dtype=torch.int64, generator=g).pin_memory() for _ in range(2)]
    trainer.load_batch(host_images[0], host_labels[0])
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        rt.control.barrier()

    ext = rt.gpu.ext

    def align_devices():
        """Flag barrier ON THE DEVICE right before the first timed event, so rank skew after the host barrier
        (python wake-up jitter) is not billed to the first step."""
        if N > 1:
            st_arena = rt.gpu.staging()[0]
            ext.barrier(st_arena.mesh, 2, torch.cuda.current_stream().cuda_stream)

    # ---------------- kernel-only: resident device batch ----------------
    sampler = ClockSampler(dev_index).start() if rt.rank == 0 else None  # before the barrier: spawning costs host time
    for _ in range(max(3, args.warmup)):  # never fewer than 3 untimed steps, whatever W says
        trainer.device_step()
    barrier()
    align_devices()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    t0.record()
    for _ in range(args.steps):
        trainer.device_step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    rt.control.barrier()
    wall = time.time() - wall0
    ms = max_over_ranks(ms)
    if opt.engine is not None:
        opt.engine.check()
    loss_val = float(trainer.loss.item())
    img_s = N * B * args.steps / (ms / 1e3)

    # ---------------- end to end: pinned host batch -> H2D -> normalise -> step -> D2H loss ----------------
    e2e = None
    if not args.no_e2e:
        fut = None
        for i in range(max(2, args.warmup // 2)):
            fut = trainer.step_async(host_images[i & 1], host_labels[i & 1])
        fut.result()
        barrier()
        align_devices()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prev = None
        for i in range(args.steps):
            cur = trainer.step_async(host_images[i & 1], host_labels[i & 1])
            if prev is not None:
                prev.result()  # D2H read of the previous step's loss, every step
            prev = cur
        last_loss = prev.result()
        e1.record()
        torch.cuda.synchronize()
        ems = max_over_ranks(e0.elapsed_time(e1))
        rt.control.barrier()
        e2e = {"value": round(N * B * args.steps / (ems / 1e3), 1), "unit": "images/sec", "ms_per_step": round(ems / args.steps, 3),
               "h2d_bytes_per_step": trainer.h2d_bytes_per_step, "d2h_bytes_per_step": trainer.d2h_bytes_per_step,
               "last_loss": round(last_loss, 4)}
    clocks = sampler.stop() if sampler is not None else None
    if args.trace:
        _trace(args, trainer, rt)

    if rt.rank == 0:
        launches = trainer.launches_per_step * args.steps
        opt_name = "SynchronousSGD" if args.optimizer == "ssgd" else "SynchronousAveraging"
        out = {
            "metric": f"ResNet-50 {opt_name} training throughput",
            "value": round(img_s, 1),
            "unit": "images/sec",
            "n_gpus": N,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3),
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": round(img_s / BASELINE_IMG_S, 3),
            "dtype": "bf16",
            "data": "synthetic",
            "impl": args.impl,
            "config": {"model": args.model, "global_batch": N * B, "per_gpu_batch": B, "image": "3x224x224",
                       "parallelism": f"dp{N}", "optimizer": f"{opt_name}(SGD lr={lr}, momentum={args.momentum})",
                       "collective": "ncclAllReduce via torch.distributed (baseline arm)" if backend == "nccl" else
                       ("split-phase sm_100a kernels over peer-mapped HBM" if opt.engine.gated else "single-kernel sm_100a collectives"),
                       "memory_format": "channels_last", "cuda_graph": trainer.graph is not None,
                       "l2_policy": "inputs+activations per step (>1 GB) exceed the 126 MB L2; no explicit flush",
                       "baseline": "5507 img/s = reference S-SGD on 16xV100 (whole cluster), BASELINE.md"},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "launches_per_step": int(trainer.launches_per_step),
            "clocks": clocks,
            "loss": round(loss_val, 4),
            "wall_s": round(wall, 3),
        }
        if trainer.graph_error:
            out["config"]["cuda_graph_error"] = trainer.graph_error[:200]
        if opt.engine is not None:
            out["config"]["mesh"] = {"mode": opt.engine.arena.mesh.mode, "multicast": bool(opt.engine.arena.has_multicast),
                                     "note": opt.engine.arena.mesh.multicast_note, "buckets": sum(len(s.layout.buckets) for s in opt.engine.sets)}
        print(json.dumps(out))
    torch.cuda.synchronize()
    return 0


if __name__ == "__main__":
    rc = main()
    sys.stdout.flush()
    os._exit(rc) if rc else sys.exit(0)
