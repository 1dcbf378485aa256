"""Multi-GPU correctness worker (run under torchrun, one rank per GPU).

Sweeps the collective kernels (all dtypes/ops/algorithms, 16 B … 64 MiB) against a host-computed reference — the
counterpart of the NCCL `simple_test` 1–90 MiB sweep (tests/cpp/integration/fake_nccl_trainer.cpp:34-63) and of
`test_torch_ops.py` (all_reduce == x·np, all_gather stack) — then checks the fused S-SGD / SMA / PairAveraging
kernels against plain PyTorch fp32 implementations of the same update.
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200.ops import collective as coll  # noqa: E402
from kungfu_b200.runtime import get_runtime  # noqa: E402
from kungfu_b200.utils import dtypes  # noqa: E402


def log(*a):
    if get_runtime().rank == 0:
        print(*a, flush=True)


def rank_tensor(n, dtype, rank, seed=0):
    g = torch.Generator(device="cuda").manual_seed(1000 * seed + rank)
    if dtype.is_floating_point:
        return (torch.randn(n, device="cuda", generator=g) * 2).to(dtype)
    return torch.randint(1, 5, (n,), device="cuda", generator=g).to(dtype)


def check_allreduce(rt):
    N = rt.size
    fails = 0
    plane = rt.gpu
    arena, off, cap = plane.staging()
    log(f"mesh mode={arena.mesh.mode} multicast={arena.has_multicast} note={arena.mesh.multicast_note}")
    cases = []
    for dtype in (torch.float32, torch.bfloat16, torch.float16, torch.int32, torch.int64, torch.uint8, torch.float64):
        for op in ("sum", "min", "max", "prod"):
            cases.append((dtype, op))
    sizes = [1, 7, 1024, 65536 + 3, (1 << 20) + 5]
    algos = ["ll", "oneshot", "twoshot"] + (["nvls"] if arena.has_multicast else [])
    for algo in algos:
        for dtype, op in cases:
            for n in sizes:
                if op == "prod" and n > 1024:
                    continue
                xs = [rank_tensor(n, dtype, r, seed=n % 97) for r in range(N)]
                if op == "prod":
                    xs = [(x.float().abs().clamp(0.5, 1.5)).to(dtype) if dtype.is_floating_point else x.clamp(1, 2) for x in xs]
                acc_t = torch.float64 if dtype.is_floating_point else torch.int64
                ref = xs[0].to(acc_t)
                for x in xs[1:]:
                    y = x.to(acc_t)
                    ref = {"sum": ref + y, "min": torch.minimum(ref, y), "max": torch.maximum(ref, y), "prod": ref * y}[op]
                mine = xs[rt.rank].clone()
                coll._run_on_comm(plane, lambda s: coll._gpu_all_reduce_(mine, dtypes.op_code(op)[0], False, stream=s, algo=algo), (mine,))
                torch.cuda.synchronize()
                arena.check()
                if dtype.is_floating_point:
                    tol = {torch.float32: 1e-5, torch.float64: 1e-12, torch.bfloat16: 2e-2, torch.float16: 2e-3}[dtype]
                    ok = torch.allclose(mine.to(torch.float64), ref, rtol=tol, atol=tol * N * 4)
                else:
                    ok = torch.equal(mine.to(torch.int64), ref.to(dtype).to(torch.int64))
                # every rank must hold bit-identical results
                same = coll.consensus(mine)
                if not (ok and same):
                    fails += 1
                    log(f"FAIL allreduce algo={algo} dtype={dtype} op={op} n={n} ok={ok} same={same}")
    # averaging + big message through the public API
    x = rank_tensor((16 << 20) + 13, torch.bfloat16, rt.rank, 5)
    ref = sum(rank_tensor((16 << 20) + 13, torch.bfloat16, r, 5).float() for r in range(N)) / N
    coll.all_reduce_(x, "avg")
    torch.cuda.synchronize()
    if not torch.allclose(x.float(), ref, rtol=2e-2, atol=2e-2):
        fails += 1
        log("FAIL avg big", (x.float() - ref).abs().max().item())
    return fails


def check_bcast_gather(rt):
    fails = 0
    N = rt.size
    for n in (3, 4099, (4 << 20) + 1):
        x = rank_tensor(n, torch.float32, rt.rank, 9)
        want = rank_tensor(n, torch.float32, N - 1, 9)
        coll.broadcast_(x, root=N - 1)
        torch.cuda.synchronize()
        if not torch.equal(x, want):
            fails += 1
            log("FAIL broadcast", n)
        y = rank_tensor(min(n, 100000), torch.bfloat16, rt.rank, 11)
        g = coll.all_gather(y)
        torch.cuda.synchronize()
        for r in range(N):
            if not torch.equal(g[r], rank_tensor(min(n, 100000), torch.bfloat16, r, 11)):
                fails += 1
                log("FAIL all_gather", n, r)
    coll.barrier()
    return fails


def mlp(dtype, seed=3):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(96, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
                               torch.nn.Linear(256, 10)).cuda().to(dtype)


def batch(rank, step, n=32):
    g = torch.Generator(device="cuda").manual_seed(77 * step + rank)
    return torch.randn(n, 96, device="cuda", generator=g), torch.randint(0, 10, (n,), device="cuda", generator=g)


def check_ssgd(rt, algo, optname, dtype, monitor=False):
    """Fused S-SGD over N ranks == single-process torch optimizer on the averaged gradient."""
    N = rt.size
    ref = mlp(torch.float32)
    m = mlp(dtype)
    with torch.no_grad():
        for a, b in zip(ref.parameters(), m.parameters()):
            a.copy_(b.float())
    mk = (lambda p: torch.optim.SGD(p, lr=0.05, momentum=0.9, weight_decay=1e-3)) if optname == "sgd" else (
        lambda p: torch.optim.AdamW(p, lr=2e-3, weight_decay=1e-2))
    o_ref = mk(ref.parameters())
    opt = kf.optimizers.SynchronousSGDOptimizer(mk(m.parameters()), named_parameters=m.named_parameters(), algo=algo,
                                                bucket_mb=0.05, monitor=monitor)
    eng = opt.engine
    fails = 0
    for step in range(6):
        # reference: average of the per-rank gradients computed with the SAME (possibly bf16) model weights
        o_ref.zero_grad()
        grads = None
        with torch.no_grad():
            snap = [p.detach().clone() for p in m.parameters()]
        for r in range(N):
            mm = mlp(dtype)
            with torch.no_grad():
                for a, b in zip(mm.parameters(), snap):
                    a.copy_(b)
            x, y = batch(r, step)
            torch.nn.functional.cross_entropy(mm(x.to(dtype)).float(), y).backward()
            gs = [p.grad.float() for p in mm.parameters()]
            grads = gs if grads is None else [a + b for a, b in zip(grads, gs)]
        for p, g in zip(ref.parameters(), grads):
            p.grad = g / N
        o_ref.step()
        opt.zero_grad()
        x, y = batch(rt.rank, step)
        torch.nn.functional.cross_entropy(m(x.to(dtype)).float(), y).backward()
        opt.step()
        torch.cuda.synchronize()
        eng.check()
        if monitor:
            sq = eng.read_sumsq()
            pre = sum((g.to(dtype).float() ** 2).sum() for g in gs)  # last rank's local grads only — checked loosely
            post = sum(((g / N) ** 2).sum() for g in grads)
            if not math.isclose(sq[1].item(), post.item(), rel_tol=5e-2):
                fails += 1
                log(f"FAIL sumsq post {sq[1].item()} vs {post.item()}")
    tol = 2e-5 if dtype == torch.float32 else 3e-2
    for a, b in zip(ref.parameters(), m.parameters()):
        if not torch.allclose(a, b.float(), atol=tol, rtol=tol):
            fails += 1
            log(f"FAIL ssgd algo={algo} opt={optname} dtype={dtype} maxdiff={(a - b.float()).abs().max().item()}")
            break
    same = all(coll.consensus(p.data) for p in m.parameters())
    if not same:
        fails += 1
        log(f"FAIL ssgd params differ across ranks algo={algo}")
    return fails


def main():
    rt = get_runtime()
    torch.cuda.set_device(rt.cuda_index())
    what = sys.argv[1:] or ["allreduce", "bcast", "ssgd"]
    fails = 0
    if "allreduce" in what:
        fails += check_allreduce(rt)
        log("allreduce sweep done, fails =", fails)
    if "bcast" in what:
        fails += check_bcast_gather(rt)
        log("bcast/gather done, fails =", fails)
    if "ssgd" in what:
        arena = rt.gpu.staging()[0]
        algos = ["twoshot"] + (["nvls"] if arena.has_multicast else [])
        for algo in algos:
            fails += check_ssgd(rt, algo, "sgd", torch.float32)
            fails += check_ssgd(rt, algo, "adam", torch.float32)
            fails += check_ssgd(rt, algo, "sgd", torch.bfloat16, monitor=True)
        log("ssgd done, fails =", fails)
    if "extra" in what:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import mp_gpu_extra

        fails += mp_gpu_extra.run(rt, log)
    tot = torch.tensor([fails], dtype=torch.int64)
    rt.control.all_reduce_cpu_(tot, "sum")
    if rt.rank == 0:
        print(json.dumps({"mp_gpu_worker": "done", "fails": int(tot.item()), "world": rt.size}), flush=True)
    torch.cuda.synchronize()
    sys.exit(1 if tot.item() else 0)


if __name__ == "__main__":
    main()
