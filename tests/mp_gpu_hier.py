"""Host-spanning GPU paths on ONE box: `KUNGFU_FORCE_GROUP_SIZE=g` splits the ranks into NVLink groups of g that behave like
separate hosts (kernels inside a group, pinned memory + control-plane subset all-reduce between the groups' first ranks).

Counterpart of the reference's multi-host runs (`kungfu-run -H a:4,b:4`, CrossAllReduceGpu staging through host memory,
srcs/cpp/src/nccl/controller.cpp). Run under torchrun / kungfu-run with 2 GPUs (g=1) or 4 GPUs (g=2).
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200 import ops  # noqa: E402
from kungfu_b200.ops import collective as coll  # noqa: E402
from kungfu_b200.runtime import get_runtime  # noqa: E402


def main():
    rt = get_runtime()
    N, me = rt.size, rt.rank
    torch.cuda.set_device(rt.cuda_index())
    plane = rt.gpu
    fails = []

    def check(ok, what):
        if not ok:
            fails.append(what)
            print(f"[rank {me}] FAIL {what}", flush=True)

    want_group = int(os.environ.get("KUNGFU_FORCE_GROUP_SIZE") or os.environ.get("KUNGFU_EXPECT_GROUP") or 1)
    check(plane.spans_hosts and plane.group_size == want_group, "the job spans NVLink groups")
    if me == 0:
        print(f"groups of {plane.group_size}: masters {coll._group_masters(plane)}", flush=True)
    else:
        coll._group_masters(plane)
    # ---- all-reduce: sizes around the chunking, dtypes, ops ----
    for dtype in (torch.float32, torch.bfloat16, torch.int32):
        for n in (1, 1000, (1 << 20) + 3, 20 * (1 << 20) + 5):
            if n > (1 << 21) and dtype != torch.float32:
                continue
            xs = [(torch.arange(n, device="cuda") % 13 + r + 1).to(dtype) for r in range(N)]
            for op in ("sum", "max", "avg"):
                if op == "avg" and not dtype.is_floating_point:
                    continue
                want = {"sum": sum(x.double() for x in xs), "max": torch.stack([x.double() for x in xs]).max(0).values,
                        "avg": sum(x.double() for x in xs) / N}[op]
                got = ops.all_reduce(xs[me], op)
                tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
                check(torch.allclose(got.double(), want, rtol=tol, atol=tol), f"all_reduce {dtype} n={n} {op}")
    # ---- broadcast / all-gather / barrier ----
    t = torch.full((12345,), float(me + 1), device="cuda")
    ops.broadcast_(t, root=N - 1)
    check(torch.all(t == float(N)).item(), "broadcast from the last rank")
    g = ops.all_gather(torch.full((5, 3), float(me), device="cuda"))
    check(g.shape == (N, 5, 3) and all(torch.all(g[r] == float(r)).item() for r in range(N)), "all_gather")
    ops.barrier()
    # ---- P2P store across groups ----
    ops.save_variable(torch.full((4096,), float(100 + me), device="cuda"), "hier-blob")
    ops.barrier()
    ok, got = ops.request_variable((me + 1) % N, "hier-blob", shape=(4096,), dtype=torch.float32, device="cuda")
    check(ok and torch.all(got == float(100 + (me + 1) % N)).item(), "request_variable across groups")
    ok, _ = ops.request_variable((me + 1) % N, "no-such-blob", shape=(4,), dtype=torch.float32, device="cuda")
    check(not ok, "missing blob reports failure")
    ops.barrier()
    # ---- S-SGD takes the generic path and still equals the single-process reference ----
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.Tanh(), torch.nn.Linear(64, 4)).cuda()
    ref = torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.Tanh(), torch.nn.Linear(64, 4)).cuda()
    ref.load_state_dict(model.state_dict())
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9), named_parameters=model.named_parameters())
    check(opt.engine is None, "fused engine disabled for host-spanning jobs")
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9)
    for step in range(4):
        gen = torch.Generator(device="cuda").manual_seed(step)
        x = torch.randn(N * 8, 32, device="cuda", generator=gen)
        y = torch.randint(0, 4, (N * 8,), device="cuda", generator=gen)
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(x[me * 8:(me + 1) * 8]), y[me * 8:(me + 1) * 8]).backward()
        opt.step()
        ropt.zero_grad()
        torch.nn.functional.cross_entropy(ref(x), y).backward()  # the mean over the global batch = the mean of the shard means
        ropt.step()
    for a, b in zip(model.parameters(), ref.parameters()):
        check(torch.allclose(a, b, atol=1e-5, rtol=1e-5), f"S-SGD vs single process, maxdiff {(a - b).abs().max().item():.2e}")
    # ---- SMA / PairAveraging generic paths run and keep replicas finite ----
    for mk in (kf.optimizers.SynchronousAveragingOptimizer, kf.optimizers.PairAveragingOptimizer):
        m2 = torch.nn.Linear(16, 4).cuda()
        o2 = mk(torch.optim.SGD(m2.parameters(), lr=0.05), named_parameters=m2.named_parameters())
        check(o2.engine is None, f"{mk.__name__}: generic path")
        for step in range(3):
            o2.zero_grad()
            m2(torch.randn(8, 16, device="cuda")).pow(2).mean().backward()
            o2.step()
        check(all(torch.isfinite(p).all().item() for p in m2.parameters()), f"{mk.__name__}: finite parameters")
        ops.barrier()
    torch.cuda.synchronize()
    print(json.dumps({"mp_gpu_hier": "done", "fails": len(fails), "world": N, "group": plane.group_size}), flush=True)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
