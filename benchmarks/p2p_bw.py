#!/usr/bin/env python
"""Peer-to-peer pull bandwidth and request rate: `request_variable` over NVLink (LDG/STG kernel vs the TMA bulk-copy kernel).

Counterpart of the reference's P2P benchmark (tests/go/cmd/kungfu-bench-p2p: request latency / rate over TCP). Every rank
publishes a blob in the symmetric arena, then pulls its right neighbour's copy; time is device time (CUDA events), max over
ranks; bandwidth = bytes / t against the measured ~770 GB/s unidirectional NVLink 5 peer-copy rate.

  torchrun --nproc-per-node 2 benchmarks/p2p_bw.py [--max-mb 256] [--out profiles/p2p_bw_2gpu.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from kungfu_b200.ops import collective as coll  # noqa: E402
from kungfu_b200.runtime import get_runtime  # noqa: E402
from kungfu_b200.utils.timing import ClockSampler, max_over_ranks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-mb", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rt = get_runtime()
    N = rt.size
    torch.cuda.set_device(rt.cuda_index())
    plane = rt.gpu
    ext = plane.ext
    max_bytes = args.max_mb << 20
    arena = plane.new_arena(2 * max_bytes + (4 << 20), "p2pbw").ensure()
    pub0, pub1, ver = arena.alloc(max_bytes, align=4096), arena.alloc(max_bytes, align=4096), arena.alloc(256)
    src = torch.full((max_bytes // 4,), float(rt.rank + 1), device="cuda")
    dst = torch.empty(max_bytes // 4, device="cuda")
    done = torch.zeros(4, dtype=torch.int32, device="cuda")
    ok = torch.zeros(4, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ext.publish(arena.mesh, 1, [pub0, pub1], ver, src.data_ptr(), max_bytes, done.data_ptr(), 64, 512, st)
    torch.cuda.synchronize()
    coll.barrier()
    target = (rt.rank + 1) % N
    sampler = ClockSampler(rt.cuda_index()).start() if rt.rank == 0 else None
    rows = []

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        coll.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return max_over_ranks(e0.elapsed_time(e1) / args.iters)

    nb = 4096
    while nb <= max_bytes:
        variants = {}
        for blocks in (32, 74, 148):
            variants[f"ldg_b{blocks}"] = lambda b=blocks: ext.pull(arena.mesh, 1, target, [pub0, pub1], ver, dst.data_ptr(), nb, ok.data_ptr(), 0, b, 512, st)
        for blocks in (37, 74, 148):
            variants[f"tma_b{blocks}"] = lambda b=blocks: ext.pull_tma(arena.mesh, 1, target, [pub0, pub1], ver, dst.data_ptr(), nb, ok.data_ptr(), 0,
                                                                        min(b, (nb + 32767) // 32768), st)
        res = {k: timed(f) for k, f in variants.items()}
        best_ldg = min(v for k, v in res.items() if k.startswith("ldg"))
        best_tma = min(v for k, v in res.items() if k.startswith("tma"))
        row = {"bytes": nb, "ldg_us": round(best_ldg * 1e3, 2), "tma_us": round(best_tma * 1e3, 2),
               "ldg_GBs": round(nb / best_ldg / 1e6, 1), "tma_GBs": round(nb / best_tma / 1e6, 1),
               "detail_us": {k: round(v * 1e3, 2) for k, v in res.items()}}
        rows.append(row)
        if rt.rank == 0:
            print(json.dumps(row), flush=True)
        nb *= 4
    assert int(ok[0].item()) == 1 and torch.all(dst == float(target + 1)), "pull returned the wrong bytes"
    if rt.rank == 0:
        small = rows[0]
        print(f"RESULT: p2p request latency {small['bytes']} B: ldg {small['ldg_us']} us, tma {small['tma_us']} us "
              f"({1e6 / min(small['ldg_us'], small['tma_us']):.0f} requests/s back to back); "
              f"peak pull bandwidth ldg {max(r['ldg_GBs'] for r in rows)} GB/s, tma {max(r['tma_GBs'] for r in rows)} GB/s", flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump({"n_gpus": N, "rows": rows, "clocks": sampler.stop() if sampler is not None else None}, f, indent=1)
    coll.barrier()


if __name__ == "__main__":
    main()
