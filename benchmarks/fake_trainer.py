#!/usr/bin/env python
"""GPU comm-only fake trainer with pluggable collective back-ends: "Img/sec, communication only".

No model, no compute: every step reduces a model-shaped list of gradient tensors (ResNet-50 / VGG16 / BERT-large shapes)
across the ranks, exactly as a data-parallel step would, and nothing else. It is the driver-independent collective metric:
the same tensor list goes through
    ours-fused   one flat arena, the engine's split-phase bucketed all-reduce kernels (what the optimizers use)
    ours-tensor  one named all-reduce per tensor through the public `all_reduce_async_` API (the reference's torch model)
    nccl-fused   the same buckets through `ncclAllReduce` (torch.distributed)
    nccl-tensor  one `ncclAllReduce` per tensor
and is reported as Img/sec = steps * 32 / t (the reference's formula), device-timed, max over ranks, plus bus GB/s.

Before timing, `simple_test` checks every back-end for message sizes 1 - 90 MiB against a host-computed sum.

Reference: tests/cpp/integration/fake_trainer.hpp:169-227 (run_experiment, "Img/sec"), collective_nccl_impl.hpp:36-93 (NCCL
back-end on device buffers), fake_nccl_trainer.cpp:34-63 (simple_test 1-90 MiB), scripts/tests/run-fake-trainer.sh.

  torchrun --nproc-per-node N benchmarks/fake_trainer.py [--model resnet50] [--dtype bf16] [--steps 50] [--out profiles/fake_trainer_N.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from kungfu_b200.benchmarks import model_sizes  # noqa: E402
from kungfu_b200.ops import collective as coll  # noqa: E402
from kungfu_b200.parallel.engine import FlatEngine  # noqa: E402
from kungfu_b200.runtime import get_runtime  # noqa: E402
from kungfu_b200.utils.timing import ClockSampler, max_over_ranks  # noqa: E402

BATCH = 32  # images a fake step stands for (fake_trainer.hpp:176)


class Backend:
    name = "?"

    def step(self) -> None:
        raise NotImplementedError

    def fill(self, value: float) -> None:
        raise NotImplementedError

    def sample(self) -> torch.Tensor:
        raise NotImplementedError


class EngineBackend(Backend):
    """Flat arena + bucketed all-reduce, `backend` = our kernels or NCCL."""

    def __init__(self, sizes, dtype, backend: str) -> None:
        self.name = ("ours" if backend == "kernels" else "nccl") + "-fused"
        self.params = [torch.nn.Parameter(torch.zeros(n, dtype=dtype, device="cuda")) for n in sizes]
        groups = [{"params": self.params, "lr": 0.0}]
        self.eng = FlatEngine(groups, [(f"t{i}", p) for i, p in enumerate(self.params)], mode="grads", opt_kind=None, average=False,
                              overlap=False, name="fake-" + backend, backend=backend)

    def fill(self, value: float) -> None:
        for s in self.eng.sets:
            s.grads.fill_(value)

    def step(self) -> None:
        self.eng.step_buckets()

    def sample(self) -> torch.Tensor:
        return self.eng.sets[0].grads[:1024].float()


class TensorBackend(Backend):
    """One all-reduce per tensor (the reference's per-tensor torch path), ours or NCCL."""

    def __init__(self, sizes, dtype, nccl: bool) -> None:
        self.name = ("nccl" if nccl else "ours") + "-tensor"
        self.nccl = nccl
        self.tensors = [torch.zeros(n, dtype=dtype, device="cuda") for n in sizes]

    def fill(self, value: float) -> None:
        for t in self.tensors:
            t.fill_(value)

    def step(self) -> None:
        if self.nccl:
            import torch.distributed as dist

            for t in self.tensors:
                dist.all_reduce(t)
        else:
            coll.all_reduce_group_async_(self.tensors, "sum").wait()  # one kernel per tensor, one stream hop for the group

    def sample(self) -> torch.Tensor:
        return self.tensors[0][:1024].float()


def simple_test(rt, log) -> int:
    """1 - 90 MiB messages through both collectives: result must be sum_r (r + 1) everywhere."""
    import torch.distributed as dist

    N = rt.size
    fails = 0
    want = float(N * (N + 1) // 2)
    for mib in (1, 2, 4, 8, 16, 32, 64, 90):
        n = (mib << 20) // 4
        for which in ("ours", "nccl"):
            x = torch.full((n,), float(rt.rank + 1), device="cuda")
            if which == "ours":
                coll.all_reduce_(x, "sum")
            elif N > 1:
                dist.all_reduce(x)
            torch.cuda.synchronize()
            ok = bool((x == want).all().item()) if N > 1 or which == "ours" else True
            if not ok:
                fails += 1
                log(f"FAIL simple_test {which} {mib} MiB: got {x[0].item()} want {want}")
    return fails


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--backends", default="ours-fused,nccl-fused,ours-tensor,nccl-tensor")
    ap.add_argument("--skip-simple-test", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rt = get_runtime()
    N = rt.size
    torch.cuda.set_device(rt.cuda_index())
    dev = torch.device("cuda", rt.cuda_index())

    def log(*a):
        if rt.rank == 0:
            print(*a, flush=True)

    if N > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.ones(1 << 16, device="cuda"))
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    sizes = model_sizes.sizes(args.model)
    total = sum(sizes)
    nbytes = total * dtype.itemsize
    log(f"model {args.model}: {len(sizes)} tensors, {total / 1e6:.2f} M elements, {nbytes / 2**20:.1f} MiB {args.dtype}; {N} ranks")
    fails = 0 if args.skip_simple_test else simple_test(rt, log)
    log(f"simple_test (1-90 MiB, ours + nccl): fails = {fails}")
    sampler = ClockSampler(rt.cuda_index()).start() if rt.rank == 0 else None
    results = []
    for name in args.backends.split(","):
        if name.endswith("-fused"):
            be = EngineBackend(sizes, dtype, "kernels" if name.startswith("ours") else "nccl")
        else:
            be = TensorBackend(sizes, dtype, nccl=name.startswith("nccl"))
        if N == 1 and name.startswith("nccl"):
            continue
        # correctness of this back-end on the model-shaped list: sum_r (r + 1)
        be.fill(float(rt.rank + 1))
        be.step()
        torch.cuda.synchronize()
        want = float(N * (N + 1) // 2)
        if not bool((be.sample() == want).all().item()):
            fails += 1
            log(f"FAIL {be.name}: reduced value {be.sample()[0].item()} != {want}")
        be.fill(0.0)  # keep values finite over many summing steps
        for _ in range(args.warmup):
            be.step()
        torch.cuda.synchronize()
        rt.control.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            be.step()
        t1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(t0.elapsed_time(t1)) / args.steps
        img_s = BATCH * N / (ms / 1e3)  # whole job: every rank "trains" BATCH images per step
        bus = 2 * (N - 1) / N * nbytes / (ms * 1e-3) / 1e9 if N > 1 else 0.0
        row = {"backend": be.name, "ms_per_step": round(ms, 4), "img_per_sec_comm_only": round(img_s, 1), "busbw_gbs": round(bus, 1),
               "frac_of_900": round(bus / 900, 3), "frac_of_770_measured_peer_copy": round(bus / 770, 3)}
        results.append(row)
        log("RESULT: " + json.dumps(row))
        del be
        torch.cuda.synchronize()
        rt.control.barrier()
    clocks = sampler.stop() if sampler is not None else None
    if rt.rank == 0:
        out = {"bench": "fake_trainer", "model": args.model, "dtype": args.dtype, "n_gpus": N, "tensors": len(sizes), "bytes": nbytes,
               "steps": args.steps, "batch_per_step": BATCH, "fails": fails, "results": results, "clocks": clocks}
        print(json.dumps(out))
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(out, f, indent=1)
    torch.cuda.synchronize()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
