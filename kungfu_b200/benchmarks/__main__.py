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
    ap.add_argument("--p2p", action="store_true", help="peer-to-peer model requests instead of all-reduce (kungfu-bench-p2p)")
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
    if a.p2p:
        return bench_p2p(rt, a, ts, nbytes)

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


def bench_p2p(rt, a, ts, nbytes) -> None:
    """Every rank publishes the model's tensors, then pulls all of them from its right neighbour each epoch. Reference:
    tests/go/cmd/kungfu-bench-p2p/kungfu-bench-p2p.go:89-119 — rate = 2 * np * model_bytes * epochs / t (request + response
    counted like the reference does). CPU tensors go through the TCP / Unix-socket blob store, CUDA tensors through the
    published double buffers in the symmetric arenas (NVLink pulls)."""
    from ..ops import p2p

    n = rt.size
    names = [f"p2p-bench-{i}" for i in range(len(ts))]
    p2p.save_variables(ts, names)
    coll.barrier()
    target = (rt.rank + 1) % n

    def epoch() -> float:
        if ts[0].is_cuda:
            torch.cuda.synchronize()
        rt.control.barrier()
        t0 = time.time()
        for t, name in zip(ts, names):
            ok, _ = p2p.request_variable(target, name, template=t)
            if not ok:
                raise RuntimeError(f"request for {name} failed")
        if ts[0].is_cuda:
            torch.cuda.synchronize()
        return time.time() - t0

    for _ in range(a.warmup_steps):
        epoch()
    d = np.array([epoch() for _ in range(a.steps)])
    coll.barrier()  # nobody leaves while a neighbour is still pulling from it
    rate = 2 * n * nbytes / d / (1 << 30)
    if rt.rank == 0:
        attrs = {"np": n, "model": a.model, "method": a.method, "p2p": True, "fuse": a.fuse, "tensors": len(ts), "bytes": nbytes,
                 "requests_per_s": round(float(len(ts) / d.mean()), 1), "ms": round(float(d.mean() * 1e3), 3)}
        print("RESULT: %.3f +-%.3f %s GiB/s" % (rate.mean(), 1.96 * rate.std(), json.dumps(attrs)), flush=True)


if __name__ == "__main__":
    main()
