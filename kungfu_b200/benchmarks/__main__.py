"""All-reduce micro-benchmark over a model's gradient list.

Reference: srcs/python/kungfu/tensorflow/v1/benchmarks/__main__.py:113-171 — all-reduce every tensor of a model (or the
fused tensor), 10 warm-up + N timed epochs, report the "equivalent data rate" `total_bytes * 4(np-1) / t` as
`RESULT: <mean> +-<1.96 sigma> {json attrs}`. We print that line AND the NCCL-tests style bus bandwidth
`2(np-1)/np * bytes / t`. Works on CPU tensors (native graph collectives) and CUDA tensors (sm_100a kernels).

    kungfu-run -np 4 python -m kungfu_b200.benchmarks --model resnet50 --method CPU [--fuse]
    torchrun --nproc-per-node 8 -m kungfu_b200.benchmarks --model resnet50 --method GPU --dtype bf16
"""
from __future__ import annotations

import argparse
import json
import time

import numpy as np
import torch

from ..ops import collective as coll
from ..runtime import get_runtime
from .model_sizes import sizes


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--method", default="CPU", choices=["CPU", "GPU"])
    ap.add_argument("--fuse", action="store_true")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--warmup-steps", type=int, default=3)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    rt = get_runtime()
    n = rt.size
    dt = torch.float32 if a.dtype == "f32" else torch.bfloat16
    dev = "cpu"
    if a.method == "GPU":
        torch.cuda.set_device(rt.cuda_index())
        dev = "cuda"
    shapes = sizes(a.model)
    if a.fuse:
        shapes = [sum(shapes)]
    ts = [torch.ones(s, dtype=dt, device=dev) for s in shapes]
    nbytes = sum(t.numel() * t.element_size() for t in ts)

    def epoch() -> float:
        if dev == "cuda":
            torch.cuda.synchronize()
        rt.control.barrier()
        t0 = time.time()
        hs = [coll.all_reduce_async_(t, f"g{i}", "sum") for i, t in enumerate(ts)]
        coll.wait_all_handles(hs)
        if dev == "cuda":
            torch.cuda.synchronize()
        return time.time() - t0

    for _ in range(a.warmup_steps):
        epoch()
    d = np.array([epoch() for _ in range(a.steps)])
    rate = nbytes * 4 * (n - 1) / d / (1 << 30)
    bus = nbytes * 2 * (n - 1) / max(n, 1) / d / 1e9
    if rt.rank == 0:
        attrs = {"np": n, "model": a.model, "method": a.method, "fuse": a.fuse, "dtype": a.dtype, "tensors": len(ts), "bytes": nbytes,
                 "busbw_gbs": round(float(bus.mean()), 3), "ms": round(float(d.mean() * 1e3), 3)}
        print("RESULT: %.3f +-%.3f %s GiB/s" % (rate.mean(), 1.96 * rate.std(), json.dumps(attrs)), flush=True)


if __name__ == "__main__":
    main()
