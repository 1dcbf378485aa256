"""Multi-GPU checks of the SMA and PairAveraging kernels against in-process fp32 simulations of all N replicas."""
import torch

import kungfu_b200.torch as kf
from kungfu_b200.ops import collective as coll


def mlp(seed, dtype=torch.float32):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(48, 96), torch.nn.Tanh(), torch.nn.Linear(96, 7)).cuda().to(dtype)


def batch(rank, step, n=16):
    g = torch.Generator(device="cuda").manual_seed(991 * step + rank)
    return torch.randn(n, 48, device="cuda", generator=g), torch.randint(0, 7, (n,), device="cuda", generator=g)


def loss_of(m, rank, step):
    x, y = batch(rank, step)
    return torch.nn.functional.cross_entropy(m(x), y)


def check_sma(rt, log, algo, overlap=True):
    N, me = rt.size, rt.rank
    alpha, lr, mom = 0.1, 0.05, 0.9
    # replicas start DIFFERENT (seed = rank) so that averaging matters
    sims = [mlp(100 + r) for r in range(N)]
    sim_opts = [torch.optim.SGD(s.parameters(), lr=lr, momentum=mom) for s in sims]
    m = mlp(100 + me)
    opt = kf.optimizers.SynchronousAveragingOptimizer(torch.optim.SGD(m.parameters(), lr=lr, momentum=mom),
                                                      named_parameters=m.named_parameters(), alpha=alpha, bucket_mb=0.01, algo=algo,
                                                      overlap=overlap)
    assert opt.engine is not None and opt.engine.mode == "sma"
    fails = 0
    for step in range(5):
        # simulation: gradients at the pre-blend weights, blend with the mean model, then the local step
        for r in range(N):
            sim_opts[r].zero_grad()
            loss_of(sims[r], r, step).backward()
        with torch.no_grad():
            means = [sum(ps) / N for ps in zip(*[[p.data.clone() for p in s.parameters()] for s in sims])]
            for s in sims:
                for p, mu in zip(s.parameters(), means):
                    p.data.mul_(1 - alpha).add_(mu, alpha=alpha)
        for r in range(N):
            sim_opts[r].step()
        opt.zero_grad()
        loss_of(m, me, step).backward()
        opt.step()
        torch.cuda.synchronize()
        opt.engine.check()
    for a, b in zip(sims[me].parameters(), m.parameters()):
        if not torch.allclose(a, b, atol=2e-5, rtol=2e-5):
            fails += 1
            log(f"FAIL sma algo={algo} overlap={overlap} maxdiff={(a - b).abs().max().item()}")
            break
    return fails


def check_pair(rt, log):
    N, me = rt.size, rt.rank
    fails = 0
    lr = 0.05
    sims = [mlp(200 + r) for r in range(N)]
    sim_opts = [torch.optim.SGD(s.parameters(), lr=lr) for s in sims]
    published = [[p.data.clone() for p in s.parameters()] for s in sims]  # what each replica last published
    m = mlp(200 + me)
    opt = kf.optimizers.PairAveragingOptimizer(torch.optim.SGD(m.parameters(), lr=lr), named_parameters=m.named_parameters(),
                                               peer_selection="roundrobin")
    eng = opt.engine
    assert eng is not None and eng.mode == "pair"
    rr = [0] * N
    # Only one rank trains at a time below, so the things every member does TOGETHER at the first step of a membership (agree on the step
    # counter, bucket-plan consensus, connect the arena, publish + barrier) happen here, in the same order on every rank; inside the loop
    # the first backward of a rank would otherwise run the plan consensus from its gradient hook while the others sit in the step-counter
    # all-reduce of `_kf_membership_changed`.
    opt._kf_membership_changed()
    opt._kf_first_publish()
    # deterministic schedule: one rank steps at a time (what it pulls is then well defined)
    for step in range(3):
        for r in range(N):
            rr[r] += 1
            target = (r + 1 + (rr[r] - 1) % (N - 1)) % N
            sim_opts[r].zero_grad()
            loss_of(sims[r], r, step).backward()
            with torch.no_grad():
                for p, o in zip(sims[r].parameters(), published[target]):
                    p.data.add_(o).mul_(0.5)
            sim_opts[r].step()
            published[r] = [p.data.clone() for p in sims[r].parameters()]
            if r == me:
                opt.zero_grad()
                loss_of(m, me, step).backward()
                opt.step()
                torch.cuda.synchronize()
            coll.barrier()
    for a, b in zip(sims[me].parameters(), m.parameters()):
        if not torch.allclose(a, b, atol=2e-5, rtol=2e-5):
            fails += 1
            log(f"FAIL pair maxdiff={(a - b).abs().max().item()}")
            break
    # concurrent gossip stress with lr = 0: models must contract towards each other, nothing may hang or tear
    for g in opt.param_groups:
        g["lr"] = 0.0
    opt._kf_selection = "random"

    def spread():
        flat = torch.cat([p.data.reshape(-1) for p in m.parameters()])
        allf = coll.all_gather(flat)
        return (allf - allf.mean(0, keepdim=True)).pow(2).mean().item()

    s0 = spread()
    for step in range(60):
        opt.zero_grad()
        loss_of(m, me, 1000 + step).backward()
        opt.step()
    torch.cuda.synchronize()
    eng.check()
    coll.barrier()
    s1 = spread()
    st = eng.arena.mesh.status()
    if not (s1 < 0.05 * s0 + 1e-12):
        fails += 1
        log(f"FAIL pair gossip did not contract: {s0} -> {s1}")
    log(f"pair gossip spread {s0:.3e} -> {s1:.3e}, torn={eng.torn_reads()} (status word {st[5]})")
    return fails


def check_pull_publish(rt, log):
    """publish -> barrier -> pull(peer) returns the peer's bytes exactly; unknown versions report 0."""
    N, me = rt.size, rt.rank
    plane = rt.gpu
    ext = plane.ext
    arena = plane.new_arena(8 << 20, "pubtest").ensure()
    n = 1 << 18
    pub0, pub1, ver = arena.alloc(n * 4), arena.alloc(n * 4), arena.alloc(256)
    done = torch.zeros(4, dtype=torch.int32, device="cuda")
    ok = torch.zeros(4, dtype=torch.int32, device="cuda")
    dst = torch.zeros(n, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    fails = 0
    target = (me + 1) % N
    ext.pull(arena.mesh, 1, target, [pub0, pub1], ver, dst.data_ptr(), n * 4, ok.data_ptr(), 0, 8, 512, st)
    torch.cuda.synchronize()
    if ok[0].item() != 0:
        fails += 1
        log("FAIL pull before publish reported a version")
    coll.barrier()  # nobody publishes before every rank has done its "nothing there yet" pull
    for v in range(1, 4):
        src = torch.full((n,), float(100 * me + v), device="cuda")
        ext.publish(arena.mesh, 1, [pub0, pub1], ver, src.data_ptr(), n * 4, done.data_ptr(), 8, 512, st)
        torch.cuda.synchronize()
        coll.barrier()
        ext.pull(arena.mesh, 1, target, [pub0, pub1], ver, dst.data_ptr(), n * 4, ok.data_ptr(), 0, 8, 512, st)
        torch.cuda.synchronize()
        if ok[0].item() != v or not torch.all(dst == float(100 * target + v)):
            fails += 1
            log(f"FAIL pull v={v} got version {ok[0].item()} value {dst[0].item()}")
        # the TMA-driven pull (cp.async.bulk pipeline) must return the same bytes, including a ragged last chunk
        for nb in (n * 4, n * 4 - 48, 4096):
            dst.zero_()
            ok.zero_()
            ext.pull_tma(arena.mesh, 1, target, [pub0, pub1], ver, dst.data_ptr(), nb, ok.data_ptr(), 0, 148, st)
            torch.cuda.synchronize()
            k = nb // 4
            if ok[0].item() != v or not torch.all(dst[:k] == float(100 * target + v)) or not torch.all(dst[k:] == 0):
                fails += 1
                log(f"FAIL pull_tma v={v} nbytes={nb} got version {ok[0].item()} value {dst[0].item()}")
        coll.barrier()
    return fails


def check_checkpoint(rt, log):
    """Fused S-SGD (bf16 params, fp32 master shards, momentum): train -> save -> {keep going | fresh model + load} must give
    bit-identical parameters (the sharded state, the master weights and the device-side step counter all survive)."""
    import os
    import tempfile

    from kungfu_b200 import checkpoint as ckpt

    me = rt.rank
    path = os.path.join(tempfile.gettempdir(), "kf_ckpt_roundtrip.pt")

    def make(seed):
        m = mlp(seed, torch.bfloat16)
        o = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9), named_parameters=m.named_parameters(),
                                                  bucket_mb=0.005)
        return m, o

    def train(m, o, steps):
        for st in steps:
            o.zero_grad()
            x, y = batch(me, st)
            torch.nn.functional.cross_entropy(m(x.to(torch.bfloat16)).float(), y).backward()
            o.step()
        torch.cuda.synchronize()

    m1, o1 = make(300)
    train(m1, o1, range(3))
    ckpt.save(path, m1, o1, step=3)  # collective: every rank calls it, rank 0 writes
    coll.barrier()
    m2, o2 = make(999)  # different initial weights: everything must come from the file
    meta = ckpt.load(path, m2, o2)
    train(m1, o1, range(3, 5))
    train(m2, o2, range(3, 5))
    fails = 0
    if meta.get("step") != 3:
        fails += 1
        log(f"FAIL checkpoint meta {meta}")
    for a, b in zip(m1.parameters(), m2.parameters()):
        if not torch.equal(a, b):
            fails += 1
            log(f"FAIL checkpoint resume differs: maxdiff={(a.float() - b.float()).abs().max().item()}")
            break
    sd1, sd2 = o1.engine.state_dict(), o2.engine.state_dict()
    for k in sd1["state"]:
        if not torch.equal(sd1["state"][k], sd2["state"][k]):
            fails += 1
            log(f"FAIL checkpoint resume: optimizer state {k} differs")
    if me == 0:
        os.remove(path)
    return fails


def check_accumulation(rt, log):
    """Gradient accumulation with the overlapped engine: two micro-batches under no_sync() + one outside == one backward on the
    summed loss; a second backward WITHOUT no_sync() must raise instead of silently dropping gradients."""
    me = rt.rank
    fails = 0

    def make():
        m = mlp(400)
        o = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(m.parameters(), lr=0.05), named_parameters=m.named_parameters(), bucket_mb=0.005)
        return m, o

    ma, oa = make()
    mb, ob = make()
    oa.zero_grad()
    with oa.no_sync():
        loss_of(ma, me, 1).backward()
        loss_of(ma, me, 2).backward()
    loss_of(ma, me, 3).backward()
    oa.step()
    ob.zero_grad()
    (loss_of(mb, me, 1) + loss_of(mb, me, 2) + loss_of(mb, me, 3)).backward()
    ob.step()
    torch.cuda.synchronize()
    for a, b in zip(ma.parameters(), mb.parameters()):
        if not torch.allclose(a, b, atol=1e-5, rtol=1e-5):
            fails += 1
            log(f"FAIL accumulation under no_sync differs: {(a - b).abs().max().item()}")
            break
    ob.zero_grad()
    loss_of(mb, me, 4).backward()
    try:
        loss_of(mb, me, 5).backward()
        fails += 1
        log("FAIL a second backward without no_sync() was accepted")
    except RuntimeError:
        pass
    ob.step()  # finish the step on every rank so that the kernel sequences stay matched
    torch.cuda.synchronize()
    return fails


def run(rt, log):
    fails = 0
    fails += check_pull_publish(rt, log)
    log("pull/publish done, fails =", fails)
    arena = rt.gpu.staging()[0]
    for algo in ["twoshot"] + (["nvls"] if arena.has_multicast else []):
        fails += check_sma(rt, log, algo, overlap=True)   # average started by zero_grad(), local blend kernel after backward
        fails += check_sma(rt, log, algo, overlap=False)  # single fused kernel after backward
    log("sma done, fails =", fails)
    fails += check_checkpoint(rt, log)
    log("checkpoint done, fails =", fails)
    fails += check_accumulation(rt, log)
    log("accumulation done, fails =", fails)
    fails += check_pair(rt, log)
    log("pair done, fails =", fails)
    return fails
