#!/usr/bin/env python
"""All-reduce bus bandwidth vs message size: our one-shot / two-shot / NVLS kernels next to NCCL.

Metric: busbw = 2(N-1)/N * bytes / t, device-timed (CUDA events), max over ranks, against the
900 GB/s-per-direction NVLink 5 roofline (measured peer copy on this pool: ~770 GB/s). Counterpart of the reference's
`kungfu-bench-allreduce` / `python -m kungfu.tensorflow.v1.benchmarks` (tests/go/cmd/kungfu-bench-allreduce/
kungfu-bench-allreduce.go:73-110; srcs/python/kungfu/tensorflow/v1/benchmarks/__main__.py:135-171), whose
"equivalent data rate" 4(np-1)*S/t is also printed in a RESULT: line for the ResNet-50 gradient set.

  torchrun --nproc-per-node N benchmarks/allreduce_bw.py [--max-mb 1024] [--dtype bf16] [--out profiles/allreduce_bw_N.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from kungfu_b200.ops import collective as coll  # noqa: E402
from kungfu_b200.runtime import get_runtime  # noqa: E402
from kungfu_b200.utils import dtypes  # noqa: E402
from kungfu_b200.utils.timing import ClockSampler, max_over_ranks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-mb", type=int, default=1024)
    ap.add_argument("--min-bytes", type=int, default=1024)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--blocks", type=str, default="16,32,64")
    ap.add_argument("--nccl", action="store_true")
    ap.add_argument("--eager", action="store_true", help="time eager launches instead of a CUDA graph of `iters` back-to-back calls")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rt = get_runtime()
    N = rt.size
    torch.cuda.set_device(rt.cuda_index())
    plane = rt.gpu
    ext = plane.ext
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    code = dtypes.dtype_code(dt)
    max_bytes = args.max_mb << 20
    arena = plane.new_arena(max_bytes + (8 << 20), "bw").ensure()
    off = arena.alloc(max_bytes, align=4096)
    buf = arena.view(off, max_bytes // dt.itemsize, dt)
    buf.fill_(0.001)
    out = torch.empty(max(1 << 20, args.min_bytes), dtype=torch.uint8, device="cuda")
    nccl = None
    if args.nccl and N > 1:
        import torch.distributed as dist

        if not dist.is_initialized() or "nccl" not in str(dist.get_backend()):
            nccl = dist.new_group(backend="nccl") if dist.is_initialized() else None
            if nccl is None:
                dist.init_process_group("nccl")
                nccl = dist.group.WORLD
        nccl_buf = torch.full((max_bytes // dt.itemsize,), 0.001, dtype=dt, device="cuda")
    st = torch.cuda.current_stream()
    sampler = ClockSampler(rt.cuda_index()).start() if rt.rank == 0 else None
    rows = []
    sizes = []
    s = args.min_bytes
    while s <= max_bytes:
        sizes.append(s)
        s *= 4
    if sizes[-1] != max_bytes:
        sizes.append(max_bytes)
    algos = [("oneshot", ext.ALGO_ONESHOT), ("twoshot", ext.ALGO_TWOSHOT)] + ([("nvls", ext.ALGO_NVLS)] if arena.has_multicast else [])
    block_opts = [int(b) for b in args.blocks.split(",")]

    ll_max = 64 << 10
    ll_off = arena.alloc(ext.allreduce_ll_slot_bytes(ll_max), align=4096)
    side = torch.cuda.Stream()

    def timeit(fn, iters, nccl_op=False):
        """Device time per call, max over ranks. Default: `iters` calls captured back to back in ONE CUDA graph, so that neither
        our kernels nor NCCL are billed for host launch overhead (what a training step inside a graph sees)."""
        for _ in range(3):
            fn(st)
        torch.cuda.synchronize()
        g = None
        if not args.eager:
            try:
                side.wait_stream(st)
                with torch.cuda.stream(side):
                    fn(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local" if nccl_op else "global"):
                    for _ in range(iters):
                        fn(side)
                torch.cuda.synchronize()
            except Exception:
                g = None
                torch.cuda.synchronize()
        rt.control.barrier()
        ext.barrier(arena.mesh, 2, st.cuda_stream)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        if g is not None:
            g.replay()
        else:
            for _ in range(iters):
                fn(st)
        t1.record()
        torch.cuda.synchronize()
        arena.check()
        return max_over_ranks(t0.elapsed_time(t1) / iters)

    for nb in sizes:
        row = {"bytes": nb}
        iters = args.iters if nb <= (64 << 20) else max(5, args.iters // 4)
        for name, a in algos:
            if name == "oneshot" and nb > (4 << 20):
                continue
            best = None
            for blocks in block_opts:
                if name == "oneshot" and nb > out.numel():
                    continue
                o = out.data_ptr() if name == "oneshot" else 0
                fn = lambda s_: ext.allreduce(arena.mesh, 0, a, off, o, nb, code, 0, 1.0, 0, 0, blocks, 512, s_.cuda_stream)
                ms = timeit(fn, iters)
                if best is None or ms < best[0]:
                    best = (ms, blocks)
            if best:
                bus = 2 * (N - 1) / N * nb / (best[0] * 1e-3) / 1e9 if N > 1 else nb / (best[0] * 1e-3) / 1e9
                row[name] = {"us": round(best[0] * 1e3, 2), "busbw_gbs": round(bus, 1), "blocks": best[1]}
        if nb <= ll_max and N > 1:
            ms = timeit(lambda s_: ext.allreduce_ll(arena.mesh, 0, buf.data_ptr(), out.data_ptr(), nb, ll_off, ll_max, code, 0, 1.0, 0, s_.cuda_stream), iters)
            row["ll"] = {"us": round(ms * 1e3, 2), "busbw_gbs": round(2 * (N - 1) / N * nb / (ms * 1e-3) / 1e9, 1)}
        if nccl is not None:
            import torch.distributed as dist

            v = nccl_buf[: nb // dt.itemsize]

            def nccl_fn(s_):
                with torch.cuda.stream(s_):
                    dist.all_reduce(v, group=nccl)

            ms = timeit(nccl_fn, iters, nccl_op=True)
            row["nccl"] = {"us": round(ms * 1e3, 2), "busbw_gbs": round(2 * (N - 1) / N * nb / (ms * 1e-3) / 1e9, 1)}
        rows.append(row)
        if rt.rank == 0:
            print(json.dumps(row), flush=True)
    clocks = sampler.stop() if sampler is not None else None
    if rt.rank == 0:
        res = {"n_gpus": N, "dtype": args.dtype, "multicast": bool(arena.has_multicast), "mesh_mode": arena.mesh.mode,
               "roofline_gbs_per_direction": 900, "measured_peer_copy_gbs": 770, "timing": "eager launches" if args.eager else
               "CUDA graph of `iters` back-to-back calls, device time / iters, max over ranks", "clocks": clocks, "rows": rows}
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(res, f, indent=1)
        print("RESULT:", json.dumps({"n_gpus": N, "peak_busbw_gbs": max((r.get(k, {}).get("busbw_gbs", 0) for r in rows for k in ("twoshot", "nvls", "oneshot")), default=0)}))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
