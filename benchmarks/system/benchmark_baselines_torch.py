"""Comparator benchmarks: the same synthetic training step through the stock data-parallel methods, next to
`benchmark_kungfu_torch.py` — what the reference's `benchmark_horovod{,_torch}.py` and `benchmark_ps.py` are for (every KungFu
number sits beside a competitor measured on the same hardware; benchmarks/system/README.md:15-71).

  --method ddp   torch DistributedDataParallel over NCCL (bucketed ring/tree all-reduce overlapped with backward — the role
                 Horovod plays in the reference) + torch.optim.SGD
  --method ps    synchronous parameter server over NCCL: gradients are REDUCED to rank 0, rank 0 applies the optimizer, the
                 parameters are BROADCAST back (the reference's TF parameter-server comparator, single server)
Same model code, dtype (bf16), memory format, batch, timing (CUDA events, max over ranks) and RESULT: line as the KungFu script.
For the engine-level baseline (same fused BN / optimizer kernels and CUDA graph, only the collective swapped) see
`bench.py --impl nccl`.

    torchrun --nproc-per-node 8 benchmarks/system/benchmark_baselines_torch.py --method ddp --model resnet50
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from kungfu_b200 import models  # noqa: E402
from kungfu_b200.models import resnet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--method", default="ddp", choices=["ddp", "ps"])
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--batch-size", type=int, default=256)
    ap.add_argument("--num-warmup-batches", type=int, default=10)
    ap.add_argument("--num-batches-per-iter", type=int, default=10)
    ap.add_argument("--num-iters", type=int, default=5)
    ap.add_argument("--momentum", type=float, default=0.9)
    a = ap.parse_args()
    rank, n, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if n > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1)
    model = resnet.to_b200(models.build(a.model), torch.bfloat16, dev)
    side = models.input_size(a.model)
    opt = torch.optim.SGD(model.parameters(), lr=0.01 * n, momentum=a.momentum)
    net = model
    if a.method == "ddp" and n > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
    x = torch.randn(a.batch_size, 3, side, side, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (a.batch_size,), device=dev)
    params = [p for p in model.parameters() if p.requires_grad]

    def step():
        opt.zero_grad(set_to_none=False)
        F.cross_entropy(net(x).float(), y).backward()
        if a.method == "ps" and n > 1:
            flat = torch._utils._flatten_dense_tensors([p.grad for p in params])
            dist.reduce(flat, 0)  # workers -> server
            if rank == 0:
                for p, g in zip(params, torch._utils._unflatten_dense_tensors(flat, params)):
                    p.grad.copy_(g / n)
                opt.step()
            w = torch._utils._flatten_dense_tensors([p.data for p in params])
            dist.broadcast(w, 0)  # server -> workers
            for p, v in zip(params, torch._utils._unflatten_dense_tensors(w, params)):
                p.data.copy_(v)
        else:
            opt.step()

    for _ in range(a.num_warmup_batches):
        step()
    rates = []
    for _ in range(a.num_iters):
        torch.cuda.synchronize()
        if n > 1:
            dist.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(a.num_batches_per_iter):
            step()
        t1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if n > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        rates.append(a.batch_size * a.num_batches_per_iter / (ms.item() / 1e3))
    if rank == 0:
        r = np.array(rates)
        print("Img/sec per GPU: %.1f +-%.1f" % (r.mean(), 1.96 * r.std()))
        print("Total img/sec on %d GPU(s): %.1f +-%.1f" % (n, n * r.mean(), n * 1.96 * r.std()))
        print("RESULT: %.1f +-%.1f %s" % (n * r.mean(), n * 1.96 * r.std(), json.dumps(
            {"np": n, "model": a.model, "method": a.method, "batch_size": a.batch_size, "dtype": "bf16"})))
    if n > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
