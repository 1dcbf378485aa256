"""CPU integration worker, run under `kungfu-run -np N` (the reference's model: every integration test is an ordinary
program run by the launcher on loopback, scripts/tests/common.sh:14-25).

Sub-commands mirror the reference's integration tests:
  collectives  ≙ tests/python/integration/test_operators.py + tests/cpp/integration/fake_agent.cpp
  p2p          ≙ tests/cpp/integration/test_p2p_apis.cpp + kungfu-test-public-apis (negative request test)
  set_tree     ≙ tests/python/integration/test_set_tree.py
  optimizers   ≙ tests/python/integration/test_optimizers.py (but with the values CHECKED against closed forms)
  slp          ≙ tests/python/integration/test_mnist_slp.py (np-invariant result of S-SGD)
  elastic      ≙ tests/python/integration/test_step_based_schedule.py / kungfu-fake-adaptive-trainer
  recovery     ≙ examples/Failure_recovery_examples (automated here; the reference has no test)
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import kungfu_b200.python as kfp  # noqa: E402
import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200 import ops  # noqa: E402
from kungfu_b200.ops import collective as coll  # noqa: E402
from kungfu_b200.runtime import get_runtime  # noqa: E402

DTYPES = [torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int32, torch.int64, torch.uint8, torch.int8, torch.int16]


def check(cond, msg):
    if not cond:
        print("CHECK FAILED:", msg, flush=True)
        sys.exit(3)


def cmd_collectives(args):
    rt = get_runtime()
    n, r = rt.size, rt.rank
    check(kfp.current_rank() == r and kfp.current_cluster_size() == n, "identity")
    check(0 <= kfp.current_local_rank() < kfp.current_local_size() <= n, "local identity")
    check(kfp.uid() != 0 or n == 1, "uid")
    check(not kfp.detached(), "detached")
    coll.barrier()
    # all-reduce: y[i] = i * sum(ranks) style exact checks on every dtype / op
    for dt in DTYPES:
        base = (torch.arange(1000) % 5 + 1).to(dt)
        x = (base * (r + 1)).to(dt) if dt != torch.uint8 else base.clone()
        exp = base.to(torch.float64) * (sum(range(1, n + 1)) if dt != torch.uint8 else n)
        y = ops.all_reduce(x, "sum")
        check(torch.allclose(y.to(torch.float64), exp, rtol=2e-2 if dt in (torch.float16, torch.bfloat16) else 1e-9), f"sum {dt}")
        check(torch.equal(ops.all_reduce(x, "max").to(torch.float64), (base * (n if dt != torch.uint8 else 1)).to(dt).to(torch.float64)), f"max {dt}")
        check(torch.equal(ops.all_reduce(x, "min").to(torch.float64), base.to(torch.float64)), f"min {dt}")
    p = ops.all_reduce(torch.full((7,), 2.0), "prod")
    check(torch.equal(p, torch.full((7,), 2.0 ** n)), "prod")
    # multi-chunk (> 1 MiB) message, 1-byte message, averaging, in-place async with names
    big = torch.arange(700001, dtype=torch.int64) * (r + 1)
    coll.all_reduce_(big, "sum")
    check(torch.equal(big, torch.arange(700001, dtype=torch.int64) * sum(range(1, n + 1))), "big int64")
    one = torch.tensor([r], dtype=torch.uint8)
    check(int(ops.all_reduce(one, "max")) == n - 1, "1 byte")
    avg = ops.all_reduce(torch.full((33,), float(r)), "avg")
    check(torch.allclose(avg, torch.full((33,), (n - 1) / 2.0)), "avg")
    ts = [torch.full((17 + i,), float(r + i)) for i in range(8)]
    hs = [kf.ops.inplace_all_reduce_async_op(t, f"t{i}") for i, t in enumerate(ts)]
    kf.ops.wait_all_handles(hs)
    for i, t in enumerate(ts):
        check(torch.allclose(t, torch.full((17 + i,), float(sum(range(n)) + n * i))), f"async {i}")
    # group all-reduce incl. None
    g = ops.group_all_reduce([torch.ones(3), None, torch.ones(2) * r])
    check(g[1] is None and torch.equal(g[0], torch.ones(3) * n), "group")
    # broadcast / all_gather / consensus
    b = torch.full((1000,), float(r + 5))
    coll.broadcast_(b, 0)
    check(torch.equal(b, torch.full((1000,), 5.0)), "broadcast root 0")
    b2 = torch.full((10,), float(r))
    coll.broadcast_(b2, n - 1)
    check(torch.equal(b2, torch.full((10,), float(n - 1))), "broadcast root n-1")
    ag = kf.ops.all_gather(torch.full((2, 3), float(r)))
    check(ag.shape == (n, 2, 3) and all(torch.equal(ag[i], torch.full((2, 3), float(i))) for i in range(n)), "all_gather")
    check(ops.consensus(torch.ones(4)) is True, "consensus same")
    check(ops.consensus(torch.full((4,), float(r))) == (n == 1), "consensus differs")
    check(rt.control.consensus(b"x" * (r + 1)) == (n == 1), "consensus length differs")
    check(kfp.all_reduce_int_max(r * 10) == (n - 1) * 10, "int max")
    # reduce / gather / local / cross via the native control plane
    c = rt.control
    if hasattr(c, "reduce_cpu"):
        red = c.reduce_cpu(torch.full((5,), float(r + 1)))
        if r == 0:
            check(torch.equal(red, torch.full((5,), float(sum(range(1, n + 1))))), "reduce")
        gat = c.gather_cpu(torch.full((3,), float(r)))
        if r == 0:
            check(all(torch.equal(gat[i], torch.full((3,), float(i))) for i in range(n)), "gather")
        lr = c.local_reduce_cpu(torch.full((4,), 1.0))
        if c.local_rank == 0:
            check(torch.equal(lr, torch.full((4,), float(c.local_size))), "local reduce")
        lb = torch.full((4,), float(r))
        c.local_broadcast_cpu_(lb)
        x = torch.full((6,), float(r + 1))
        c.cross_all_reduce_cpu_(x)
        check(torch.equal(x, torch.full((6,), float(r + 1))) or c.host_count() > 1, "cross all-reduce on one host is identity")
        lat = ops.get_peer_latencies()
        check(lat.shape == (n,) and float(lat[r]) == 0.0 and all(float(v) >= 0 for v in lat), "latencies")
        _ = ops.egress_rates()
        m = ops.monitored_all_reduce(torch.ones(1 << 18, dtype=torch.int64))
        check(torch.equal(m, torch.full((1 << 18,), n, dtype=torch.int64)), "monitored")
        kfp.calc_stats(); kfp.log_stats(); kfp.print_strategy_stats()
        check(kfp.check_interference() is False, "first interference check takes the reference window")
    # reference names kept for scripts that use them: group_(hierarchical_)nccl_all_reduce, save_model / request_model, AdaSGDHook
    g2 = ops.group_nccl_all_reduce([torch.full((5,), float(r)), None, torch.full((70000,), 1.0)])
    check(g2[1] is None and torch.equal(g2[0], torch.full((5,), float(sum(range(n))))) and torch.equal(g2[2], torch.full((70000,), float(n))),
          "group_nccl_all_reduce")
    g3 = ops.group_hierarchical_nccl_all_reduce([torch.ones(4) * (r + 1)], op="max")
    check(torch.equal(g3[0], torch.full((4,), float(n))), "group_hierarchical_nccl_all_reduce")
    model = [torch.full((3, 2), float(r)), torch.full((7,), float(10 + r))]
    ops.save_model(model)
    coll.barrier()
    ok, got = ops.request_model((r + 1) % n, model)
    t = (r + 1) % n
    check(ok and torch.equal(got[0], torch.full((3, 2), float(t))) and torch.equal(got[1], torch.full((7,), float(10 + t))), "save_model / request_model")
    w = torch.full((6,), float(r))
    hook = kf.optimizers.AdaSGDHook(change_step=2, tensors=[w])
    fired = [hook.after_run(s) for s in range(4)]
    check(fired == [False, False, True, False] and torch.equal(w, torch.zeros(6)), "AdaSGDHook broadcasts once at the change step")
    peers = {kf.optimizers.get_random_peer(n, r) for _ in range(200)}
    check(r not in peers or n == 1, "get_random_peer never returns self")
    coll.barrier()
    print(json.dumps({"ok": True, "rank": r, "size": n}), flush=True)


def cmd_p2p(args):
    rt = get_runtime()
    n, r = rt.size, rt.rank
    me = torch.full((1024,), float(r + 1))
    ops.save_variable(me, "model")
    for v in range(3):
        ops.save_variable(me * (v + 1), "vmodel", version=v)
    coll.barrier()
    tgt = (r + 1) % n
    ok, got = ops.request_variable(tgt, "model", (1024,), torch.float32)
    check(ok and torch.equal(got, torch.full((1024,), float(tgt + 1))), "request")
    ok, got = ops.request_variable(tgt, "vmodel", (1024,), torch.float32, version=2)
    check(ok and torch.equal(got, torch.full((1024,), float(tgt + 1) * 3)), "versioned request")
    ok, _ = ops.request_variable(tgt, "no-such-name", (4,), torch.float32)
    check(ok is False, "unknown name must fail softly")
    ok, _ = ops.request_variable(tgt, "vmodel", (1024,), torch.float32, version=99)
    check(ok is False, "unknown version must fail softly")
    # window of 3 versions: version 0 is gone after saving version 3
    ops.save_variable(me, "vmodel", version=3)
    coll.barrier()
    ok, _ = ops.request_variable(tgt, "vmodel", (1024,), torch.float32, version=0)
    check(ok is False, "versioned store keeps a sliding window of 3")
    coll.barrier()
    # legacy fused-model ops: synchronous averaging pulls now; the asynchronous form averages in what it prefetched during the
    # previous call (first call: nothing yet) and never waits for a transfer
    if n > 1:
        w = [torch.full((8,), float(r)), torch.full((3, 3), float(r))]
        ma = ops.ModelAveraging(w, strategy="roundrobin", name="legacy")
        ma.save()
        coll.barrier()
        check(ma() is True, "model_averaging pulled a model")
        peer = (r + 1) % n if r + 1 != r else r
        first = [q for q in range(n) if q != r][0]
        check(bool(torch.all(w[0] == 0.5 * (r + first))) and bool(torch.all(w[1] == 0.5 * (r + first))), "model_averaging averaged with the selected peer")
        coll.barrier()
        v = [torch.full((16,), float(10 * r))]
        ama = ops.AsyncModelAveraging(v, strategy="roundrobin", name="legacy-async")
        ama.save()
        coll.barrier()
        check(ama() is False, "async_model_averaging: the first call only starts the prefetch")
        check(ama() is True, "async_model_averaging: the second call averages the prefetched model in")
        check(bool(torch.all(v[0] == 0.5 * (10 * r + 10 * first))), "async_model_averaging value")
        coll.barrier()
    print(json.dumps({"ok": True}), flush=True)


def cmd_set_tree(args):
    import random

    rt = get_runtime()
    n, r = rt.size, rt.rank
    rng = random.Random(1234)  # same trees on every rank
    times = []
    for it in range(16):
        father = [0] * n
        order = list(range(n))
        rng.shuffle(order)
        father[order[0]] = order[0]
        for i in range(1, n):
            father[order[i]] = order[rng.randrange(i)]
        t0 = time.time()
        ops.set_tree(father)
        times.append(time.time() - t0)
        x = ops.all_reduce(torch.full((4096,), float(r + 1)))
        check(torch.equal(x, torch.full((4096,), float(sum(range(1, n + 1))))), f"all-reduce after set_tree {father}")
    # latency-driven MST -> set_tree
    lat = ops.get_peer_latencies()
    father = ops.minimum_spanning_tree_father(lat)
    check(len(father) == n and father[0] == 0, "mst father")
    ops.set_tree(father)
    edges = ops.global_minimum_spanning_tree(lat)
    check(edges.shape == (max(0, n - 1), 2), "mst edges")
    mask = ops.get_neighbour_mask(edges)
    check(mask.shape == (n,), "neighbour mask")
    x = ops.all_reduce(torch.ones(10))
    check(torch.equal(x, torch.full((10,), float(n))), "all-reduce on the MST")
    # gather -> transform on rank 0 -> broadcast
    y = ops.all_gather_transform(torch.full((6,), float(r + 1)), lambda s: s.max(0).values * 10)
    check(torch.equal(y, torch.full((6,), float(10 * n))), "all_gather_transform")
    # subset all-reduce: independent all-reduce per tree of a forest; isolated peers keep their value and may skip the call
    if n >= 2:
        pairs = [2 * (i // 2) if 2 * (i // 2) + 1 < n else i for i in range(n)]  # {0,1} {2,3} ... (+ a singleton if n is odd)
        mine = [i for i in range(n) if pairs[i] == pairs[r]]
        for dt in (torch.float32, torch.int64, torch.bfloat16):
            y = ops.subset_all_reduce(torch.full((1000,), r + 1).to(dt), pairs)
            check(torch.equal(y, torch.full((1000,), sum(i + 1 for i in mine)).to(dt)), f"subset all-reduce {pairs} {dt}")
        y = ops.subset_all_reduce(torch.full((8,), float(r + 1)), pairs, op="avg")
        check(torch.allclose(y, torch.full((8,), sum(i + 1 for i in mine) / len(mine))), "subset all-reduce avg")
        y = ops.subset_all_reduce(torch.full((8,), float(r)), pairs, op="max")
        check(torch.equal(y, torch.full((8,), float(max(mine)))), "subset all-reduce max")
        # one star over the even ranks (the shape of the cross-host stage); odd ranks are not part of it and do not call
        evens = [i for i in range(n) if i % 2 == 0]
        star = [0 if i % 2 == 0 else i for i in range(n)]
        if r % 2 == 0:
            y = ops.subset_all_reduce(torch.full((70000,), float(r + 1)), star)
            check(torch.equal(y, torch.full((70000,), float(sum(i + 1 for i in evens)))), "subset all-reduce over a star of masters")
        ops.barrier()
    if r == 0:
        print(f"set_tree latency: mean {1e3 * sum(times) / len(times):.2f}ms min {1e3 * min(times):.2f}ms max {1e3 * max(times):.2f}ms", flush=True)
    print(json.dumps({"ok": True}), flush=True)


def _slp():
    torch.manual_seed(0)
    m = torch.nn.Linear(20, 5)
    with torch.no_grad():
        m.weight.zero_()
        m.bias.zero_()
    return m


def _data(total=240):
    g = torch.Generator().manual_seed(42)
    x = torch.randn(total, 20, generator=g)
    w = torch.randn(20, 5, generator=g)
    y = (x @ w).argmax(1)
    return x, y


def cmd_optimizers(args):
    """Closed-form checks of S-SGD / SMA / PairAveraging on CPU tensors (the reference only smoke-runs them)."""
    rt = get_runtime()
    n, r = rt.size, rt.rank
    lr = 0.1
    # --- analytic S-SGD: f(x) = x^2 on every rank => x_k = (1 - 2 lr)^k x_0 (≙ v1/examples/__main__.py:26-46)
    x = torch.nn.Parameter(torch.tensor([1.0]))
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD([x], lr=lr), named_parameters=[("x", x)])
    for k in range(5):
        opt.zero_grad()
        (x * x).sum().backward()
        opt.step()
    check(abs(float(x) - (1 - 2 * lr) ** 5) < 1e-6, f"analytic ssgd {float(x)}")
    # op='sum' (the reference torch binding): gradient is n times larger
    x = torch.nn.Parameter(torch.tensor([1.0]))
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD([x], lr=lr / n), named_parameters=[("x", x)], op="sum")
    opt.zero_grad(); (x * x).sum().backward(); opt.step()
    check(abs(float(x) - (1 - 2 * lr)) < 1e-6, "sum semantics")
    # --- S-SGD with rank-dependent gradients == averaged gradient
    w = torch.nn.Parameter(torch.zeros(4))
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD([w], lr=1.0), named_parameters=[("w", w)])
    opt.zero_grad(); (w * float(r + 1)).sum().backward(); opt.step()
    check(torch.allclose(w.data, torch.full((4,), -(n + 1) / 2.0)), "ssgd averaged gradient")
    # --- SMA: v <- (1-a) v + a mean(v), then local gradient step
    a = 0.1
    v = torch.nn.Parameter(torch.full((3,), float(r)))
    sma = kf.optimizers.SynchronousAveragingOptimizer(torch.optim.SGD([v], lr=0.5), named_parameters=[("v", v)], alpha=a)
    sma.zero_grad(); (v * 2.0).sum().backward(); sma.step()
    mean = (n - 1) / 2.0
    check(torch.allclose(v.data, torch.full((3,), (1 - a) * r + a * mean - 0.5 * 2.0)), f"sma closed form {v.data}")
    # --- PairAveraging, deterministic round-robin, one rank at a time
    p = torch.nn.Parameter(torch.full((2,), float(10 * (r + 1))))
    pa = kf.optimizers.PairAveragingOptimizer(torch.optim.SGD([p], lr=1.0), named_parameters=[("p", p)], peer_selection="roundrobin")
    sim = [float(10 * (i + 1)) for i in range(n)]
    pub = list(sim)
    for k in range(n):
        if k == r:
            pa.zero_grad(); p.sum().backward(); pa.step()
        elif n > 1 and not pa._kf_published:
            # the first step of a membership is collective (publish, agree on the step counter, barrier): ranks that do not step
            # yet take part in it the way step() does
            pa._kf_membership_changed()
            rt.control.save(pa._kf_model_name, pa._kf_flat_cpu()); rt.control.barrier(); pa._kf_published = True
        if n > 1:
            tgt = (k + 1) % n
            sim[k] = 0.5 * (sim[k] + pub[tgt]) - 1.0
        else:
            sim[k] = sim[k] - 1.0
        pub[k] = sim[k]
        coll.barrier()
    check(abs(float(p[0]) - sim[r]) < 1e-5, f"pair averaging {float(p[0])} vs {sim[r]}")
    # --- monitors
    q = torch.nn.Parameter(torch.zeros(8))
    gns = kf.optimizers.MonitorGradientNoiseScaleOptimizer(torch.optim.SGD([q], lr=0.01), named_parameters=[("q", q)], device_batch_size=4)
    for k in range(3):
        gns.zero_grad(); ((q - float(r + k)) ** 2).sum().backward(); gns.step()
    _ = gns.noise_scale()
    var = kf.optimizers.MonitorGradientVarianceOptimizer(torch.optim.SGD([q], lr=0.0), named_parameters=[("q", q)])
    var.zero_grad(); (q * float(r)).sum().backward(); var.step()
    # grads: g_r = r per element; variance over ranks per element = mean(r^2) - mean(r)^2, summed over 8 elements
    ev = 8 * (sum(i * i for i in range(n)) / n - ((n - 1) / 2.0) ** 2)
    check(abs(var.variance() - ev) < 1e-4, f"variance {var.variance()} vs {ev}")
    ada = kf.optimizers.AdaptiveSGDOptimizer(torch.optim.SGD([q], lr=0.01), named_parameters=[("q", q)], change_step=2)
    for k in range(4):
        ada.zero_grad(); (q * q).sum().backward(); ada.step()
    check(ops.consensus(q.data), "after the switch to S-SGD every replica holds the same model")
    print(json.dumps({"ok": True}), flush=True)


def cmd_slp(args):
    """S-SGD over np workers with global batch B ≡ one worker with batch B: the final weights must not depend on np."""
    rt = get_runtime()
    n, r = rt.size, rt.rank
    m = _slp()
    x, y = _data()
    B = args.batch
    check(B % n == 0, "global batch must divide evenly")
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(m.parameters(), lr=0.1), named_parameters=m.named_parameters())
    kf.broadcast_parameters(m.state_dict())
    lb = B // n
    for epoch in range(2):
        for s in range(0, len(x) - B + 1, B):
            xb, yb = x[s + r * lb:s + (r + 1) * lb], y[s + r * lb:s + (r + 1) * lb]
            opt.zero_grad()
            torch.nn.functional.cross_entropy(m(xb), yb).backward()
            opt.step()
    acc = float((m(x).argmax(1) == y).float().mean())
    digest = float(m.weight.double().abs().sum() + m.bias.double().abs().sum())
    if r == 0:
        print(json.dumps({"slp": True, "np": n, "batch": B, "acc": round(acc, 6), "digest": round(digest, 6)}), flush=True)


def cmd_elastic(args):
    """Step-schedule driven resize under `kungfu-run -w`: sizes follow the schedule, state survives resizes."""
    from kungfu_b200.elastic import KungFuElasticTrainHook

    rt = get_runtime()
    m = torch.nn.Linear(4, 2)
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(m.parameters(), lr=0.01), named_parameters=m.named_parameters())
    hook = KungFuElasticTrainHook(args.schedule, args.max_step, m, opt)
    hook.begin()
    seen = []
    while True:
        # every member applies the same number of steps: value == step count on all replicas
        opt.zero_grad()
        (m(torch.ones(3, 4)).sum() * 0).backward()
        opt.step()
        with torch.no_grad():
            m.bias.add_(1.0)  # "training progress" that newcomers must inherit
        seen.append((hook.state.step, kfp.current_cluster_size()))
        check(ops.consensus(m.bias.data), f"replicas diverged at step {hook.state.step}")
        want = hook.schedule(hook.state.step)
        check(kfp.current_cluster_size() == want, f"step {hook.state.step}: size {kfp.current_cluster_size()} != schedule {want}")
        if not hook.after_step():
            break
    if kfp.detached():
        print(json.dumps({"detached": True, "step": hook.state.step}), flush=True)
        if os.environ.get("KF_DEBUG_EXIT"):
            import faulthandler

            faulthandler.dump_traceback_later(8, exit=False)
        return
    if kfp.current_rank() == 0:
        print(json.dumps({"elastic": True, "final_step": hook.state.step, "bias": float(m.bias[0]), "sizes": sorted(set(s for _, s in seen)),
                          "resizes": len(hook.profiler.records)}), flush=True)
        print(hook.profiler.report(), flush=True)


def cmd_multihost(args):
    """Semantics that only show with more than one host (run on loopback aliases 127.0.0.1, 127.0.0.2, ...): local rank / size by
    address, host count, local reduce / broadcast inside a host, cross all-reduce over the hosts' first ranks only."""
    rt = get_runtime()
    n, r = rt.size, rt.rank
    c = rt.control
    spec = os.environ["KUNGFU_SELF_SPEC"]
    hosts = [p.rsplit(":", 1)[0] for p in os.environ["KUNGFU_INIT_PEERS"].split(",")]
    mine = spec.rsplit(":", 1)[0]
    same = [i for i, h in enumerate(hosts) if h == mine]
    check(c.local_size == len(same) and c.local_rank == same.index(r), f"local identity {c.local_rank}/{c.local_size} vs {same}")
    check(c.host_count() == len(set(hosts)) and c.host_count() > 1, f"host count {c.host_count()}")
    masters = sorted({hosts.index(h) for h in set(hosts)})
    # local reduce: the host's first rank gets the sum over its host
    lr = c.local_reduce_cpu(torch.full((5,), float(r + 1)))
    if c.local_rank == 0:
        check(torch.equal(lr, torch.full((5,), float(sum(i + 1 for i in same)))), "local reduce sums the host")
    lb = torch.full((4,), float(r))
    c.local_broadcast_cpu_(lb)
    check(torch.equal(lb, torch.full((4,), float(same[0]))), "local broadcast from the host's first rank")
    x = torch.full((70000,), float(r + 1))
    c.cross_all_reduce_cpu_(x)
    want = float(sum(m + 1 for m in masters)) if r in masters else float(r + 1)
    check(torch.equal(x, torch.full((70000,), want)), f"cross all-reduce over the masters {masters}")
    # the hierarchical composition equals the flat all-reduce
    y = c.local_reduce_cpu(torch.full((9,), float(r + 1)))
    if c.local_rank != 0:
        y = torch.zeros(9)
    c.cross_all_reduce_cpu_(y)
    c.local_broadcast_cpu_(y)
    check(torch.equal(y, ops.all_reduce(torch.full((9,), float(r + 1)))), "local reduce -> cross all-reduce -> local broadcast == all-reduce")
    hs = ops.group_hierarchical_nccl_all_reduce([torch.full((9,), float(r + 1)), None, torch.arange(6, dtype=torch.float32) * (r + 1)], "sum")
    tot = float(n * (n + 1) // 2)
    check(hs[1] is None and torch.equal(hs[0], torch.full((9,), tot)) and torch.equal(hs[2], torch.arange(6, dtype=torch.float32) * tot),
          "group_hierarchical_nccl_all_reduce (local reduce -> cross all-reduce -> local broadcast per tensor) == flat all-reduce")
    # the shape of the GPU hierarchical all-reduce: only the hosts' first ranks take part in the cross-host step, then everybody
    # joins an automatically named op — op numbering is per kind, so the extra op on the masters must not shift the others
    for it in range(3):
        z = torch.full((5,), float(r + 1))
        if c.local_rank == 0:
            c.cross_all_reduce_cpu_(z)
        c.local_broadcast_cpu_(z)
        check(torch.equal(z, torch.full((5,), float(sum(m + 1 for m in masters)))), f"masters-only cross step then local broadcast ({it})")
        check(torch.equal(ops.all_reduce(torch.ones(3)), torch.full((3,), float(n))), f"auto-named all-reduce stays aligned ({it})")
    lat = ops.get_peer_latencies()
    check(lat.shape == (n,), "latencies across hosts")
    ops.barrier()
    print(json.dumps({"ok": True, "rank": r, "hosts": c.host_count(), "local": [c.local_rank, c.local_size]}), flush=True)


def cmd_follow(args):
    """Workers that only FOLLOW the config server (somebody else rewrites it, e.g. `kungfu_b200.cmd.config_client --random`):
    every step ends with `resize()`; state must stay consistent through whatever sequence of sizes arrives."""
    from kungfu_b200.elastic import ElasticState, sync_after_resize

    m = torch.nn.Linear(4, 2)
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(m.parameters(), lr=0.01), named_parameters=m.named_parameters())
    state = sync_after_resize(ElasticState(), m, opt)
    sizes, resizes = set(), 0
    while state.step < args.max_step:
        opt.zero_grad()
        (m(torch.ones(3, 4)).sum() * 0).backward()
        opt.step()
        with torch.no_grad():
            m.bias.add_(1.0)
        state.step += 1
        sizes.add(kfp.current_cluster_size())
        check(ops.consensus(m.bias.data), f"replicas diverged at step {state.step}")
        time.sleep(0.05)
        changed, det = kfp.resize(None)
        if det:
            print(json.dumps({"detached": True, "step": state.step}), flush=True)
            return
        if changed:
            resizes += 1
            state = sync_after_resize(state, m, opt)
    # every member did exactly max_step increments of the bias, wherever it joined
    check(abs(float(m.bias[0]) - float(m.bias.detach()[0])) == 0 and state.step == args.max_step, "final step")
    if kfp.current_rank() == 0:
        print(json.dumps({"followed": True, "final_step": state.step, "sizes": sorted(sizes), "resizes": resizes,
                          "final_size": kfp.current_cluster_size()}), flush=True)


def cmd_recovery(args):
    """One rank stalls inside a batch once; the launcher's failure detector must restart the job, which resumes from the
    checkpoint with the remaining epochs."""
    from kungfu_b200 import checkpoint
    from kungfu_b200.cmd import monitor_batch_begin, monitor_batch_end, monitor_epoch_end, monitor_train_end

    rt = get_runtime()
    r = rt.rank
    m = torch.nn.Linear(2, 1)
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    ck = os.path.join(args.dir, "ckpt.pt")
    meta = checkpoint.load(ck, m, opt) if args.restart else {"epoch": 0, "found": False}
    start_epoch = meta["epoch"]
    marker = os.path.join(args.dir, "crashed")
    for e in range(args.n_epochs):
        for b in range(3):
            monitor_batch_begin()
            if r == 1 and not args.restart and e == 1 and b == 1 and not os.path.exists(marker):
                open(marker, "w").write("x")
                time.sleep(3600)  # a hung worker: never reports `end`
            x = ops.all_reduce(torch.ones(1))
            monitor_batch_end()
        monitor_epoch_end()
        checkpoint.save(ck, m, opt, epoch=start_epoch + e + 1)
    monitor_train_end()
    if r == 0:
        print(json.dumps({"recovered": bool(args.restart), "epochs_run_now": args.n_epochs, "start_epoch": start_epoch}), flush=True)
    time.sleep(0.5)


def cmd_metrics(args):
    """Prometheus-style endpoint of every peer (reference: monitor/server.go, http://<peer>:<port+10000>/metrics)."""
    from kungfu_b200 import _kfrt

    rt = get_runtime()
    n, r = rt.size, rt.rank
    for _ in range(3):
        ops.all_reduce(torch.ones(1 << 16))
    ops.barrier()
    port = int(os.environ["KUNGFU_SELF_SPEC"].rsplit(":", 1)[1]) + 10000
    status, body = _kfrt.http_request("GET", f"http://127.0.0.1:{port}/metrics", "")
    check(status == 200, f"metrics status {status}")
    body = body.decode() if isinstance(body, (bytes, bytearray)) else body
    lines = [ln for ln in body.splitlines() if ln]
    eg = [ln for ln in lines if ln.startswith("egress_total_bytes{peer=")]
    check(n == 1 or len(eg) >= 1, f"egress lines: {lines[:4]}")
    check(n == 1 or any(int(ln.rsplit(" ", 1)[1]) >= (1 << 16) for ln in eg), "egress counters grew with the all-reduces")
    check(n == 1 or any(ln.startswith("ingress_total_bytes{peer=") for ln in lines), "ingress lines")
    from kungfu_b200 import monitor

    parsed = monitor.fetch_metrics()
    check(n == 1 or (len(parsed.get("egress_total_bytes", {})) >= 1 and max(parsed["egress_total_bytes"].values()) >= (1 << 16)),
          f"monitor.fetch_metrics: {parsed}")
    check(monitor.metrics_url("10.0.0.5:10003") == "http://10.0.0.5:20003/metrics", "metrics_url")
    # application metrics: cluster gauges from the runtime, gauges / counters published from Python
    check(parsed.get("kungfu_cluster_size", {}).get("") == float(n) and parsed.get("kungfu_rank", {}).get("") == float(r), f"cluster gauges {parsed}")
    monitor.set_gauge("kungfu_gradient_noise_scale", 12.5)
    monitor.add_counter('gpu_collective_bytes_total{op="all_reduce"}', 1000)
    monitor.add_counter('gpu_collective_bytes_total{op="all_reduce"}', 24)
    again = monitor.fetch_metrics()
    check(again["kungfu_gradient_noise_scale"][""] == 12.5 and again['gpu_collective_bytes_total{op="all_reduce"}'][""] == 1024.0, f"published metrics {again}")
    status, _ = _kfrt.http_request("GET", f"http://127.0.0.1:{port}/nope", "")
    check(status == 404, "unknown path is 404")
    rates = ops.egress_rates()
    check(len(rates) == n, "egress_rates has one entry per peer")
    ops.barrier()
    print(json.dumps({"ok": True, "metrics_lines": len(lines)}), flush=True)


def cmd_stress(args):
    """Randomised soak of the host plane: the same pseudo-random schedule of operations on every rank (sizes from 1 element to
    several chunks, all dtypes / ops, multi-threaded async collectives, P2P traffic in between), every result checked."""
    import random
    import threading

    rt = get_runtime()
    n, r = rt.size, rt.rank
    rng = random.Random(20240921)
    iters = int(args.extra[0]) if args.extra else 120
    sizes = [1, 3, 255, 4096, 65537, 262144, 300001, (1 << 20) + 17, 3 * (1 << 20) + 5]
    dts = [torch.float32, torch.float64, torch.int32, torch.int64, torch.bfloat16, torch.float16, torch.uint8]
    ops.save_variable(torch.full((1024,), float(r)), "soak")
    ops.barrier()
    for it in range(iters):
        kind = rng.choice(["allreduce", "allreduce", "allreduce", "bcast", "gather", "threads", "p2p", "consensus", "barrier", "subset", "mixed"])
        if kind == "allreduce":
            dt, op, m = rng.choice(dts), rng.choice(["sum", "min", "max"]), rng.choice(sizes)
            base = (torch.arange(m) % 7 + 1)
            x = (base * ((r % 3) + 1)).to(dt)
            want = {"sum": base * sum((q % 3) + 1 for q in range(n)), "min": base * min((q % 3) + 1 for q in range(n)),
                    "max": base * max((q % 3) + 1 for q in range(n))}[op].to(dt)
            got = ops.all_reduce(x, op, name=f"soak{it}")
            check(torch.equal(got, want), f"it {it}: all-reduce {dt} {op} n={m}")
        elif kind == "bcast":
            root, m = rng.randrange(n), rng.choice(sizes)
            x = torch.full((m,), float(r + 1))
            coll.broadcast_(x, root)
            check(bool(torch.all(x == float(root + 1))), f"it {it}: broadcast from {root} n={m}")
        elif kind == "gather":
            m = rng.choice(sizes[:6])
            g = ops.all_gather(torch.full((m,), r, dtype=torch.int32))
            check(g.shape == (n, m) and all(int(g[q][0]) == q and int(g[q][-1]) == q for q in range(n)), f"it {it}: all_gather n={m}")
        elif kind == "threads":
            # several named collectives in flight at once from different threads (the reference's async ops do this)
            k = rng.randrange(2, 6)
            ms = [rng.choice(sizes[:8]) for _ in range(k)]
            outs = [None] * k

            def run(i):
                outs[i] = ops.all_reduce(torch.full((ms[i],), float(i + 1)), "sum", name=f"thr{it}.{i}")

            ths = [threading.Thread(target=run, args=(i,)) for i in range(k)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            for i in range(k):
                check(outs[i] is not None and bool(torch.all(outs[i] == float((i + 1) * n))), f"it {it}: threaded all-reduce {i}")
        elif kind == "mixed":
            # different kinds of collectives at the same time from different threads (all-reduce, broadcast, all-gather, P2P pulls)
            root, m1, m2 = rng.randrange(n), rng.choice(sizes[:7]), rng.choice(sizes[:6])
            bad = []

            def t_ar():
                y = ops.all_reduce(torch.full((m1,), 2.0), name=f"mix{it}.ar")
                bad.append("ar") if not bool(torch.all(y == 2.0 * n)) else None

            def t_bc():
                x = torch.full((m2,), float(r))
                coll.broadcast_(x, root, name=f"mix{it}.bc")
                bad.append("bc") if not bool(torch.all(x == float(root))) else None

            def t_ag():
                g = ops.all_gather(torch.full((257,), float(r)), name=f"mix{it}.ag")
                bad.append("ag") if not all(bool(torch.all(g[q] == float(q))) for q in range(n)) else None

            def t_pp():
                ok, v = ops.request_variable((r + 1) % n, "soak", shape=(1024,), dtype=torch.float32)
                bad.append("p2p") if not (ok and bool(torch.all(v == float((r + 1) % n)))) else None

            ths = [threading.Thread(target=f) for f in (t_ar, t_bc, t_ag, t_pp)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            check(not bad, f"it {it}: mixed concurrent collectives {bad}")
        elif kind == "p2p":
            tgt = rng.randrange(n)
            ok, v = ops.request_variable(tgt, "soak", shape=(1024,), dtype=torch.float32)
            check(ok and bool(torch.all(v == float(tgt))), f"it {it}: request from {tgt}")
        elif kind == "consensus":
            same = rng.random() < 0.5
            check(ops.consensus(torch.full((9,), 1.0 if same else float(r))) == (same or n == 1), f"it {it}: consensus")
        elif kind == "subset" and n >= 2:
            k = rng.randrange(2, n + 1)
            members = sorted(rng.sample(range(n), k))
            forest = list(range(n))
            for q in members:
                forest[q] = members[0]
            m = rng.choice(sizes[:7])  # drawn by everyone: the schedule must stay identical on all ranks
            if r in members:
                got = ops.subset_all_reduce(torch.full((m,), float(r + 1)), forest, name=f"sub{it}")
                check(bool(torch.all(got == float(sum(q + 1 for q in members)))), f"it {it}: subset {members}")
        else:
            ops.barrier()
    ops.barrier()
    print(json.dumps({"ok": True, "iters": iters}), flush=True)


def cmd_initializer(args):
    """BroadcastGlobalVariablesOp / Callback / optimizer-state broadcast (reference: initializer.py + torch broadcast_parameters),
    scope tracing and the max-over-ranks timing helper."""
    from kungfu_b200 import initializer
    from kungfu_b200.utils import timing, trace

    rt = get_runtime()
    n, r = rt.size, rt.rank
    torch.manual_seed(100 + r)  # replicas start different
    m = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 2))
    opt = torch.optim.SGD(m.parameters(), lr=0.1 * (r + 1), momentum=0.9)
    m(torch.randn(4, 6)).sum().backward()
    opt.step()  # momentum buffers exist and differ
    with trace.trace_scope("init.broadcast"):
        initializer.BroadcastGlobalVariablesOp(m, opt)
    flat = torch.cat([v.detach().reshape(-1).float() for v in m.state_dict().values()])
    check(ops.consensus(flat), "parameters and buffers identical after BroadcastGlobalVariablesOp")
    bufs = torch.cat([opt.state[p]["momentum_buffer"].reshape(-1) for p in m.parameters()])
    check(ops.consensus(bufs), "optimizer state identical after the broadcast")
    check(opt.param_groups[0]["lr"] == 0.1, f"hyper-parameters follow the root: lr={opt.param_groups[0]['lr']}")
    # a worker that joins a running job has NO optimizer state yet (torch creates it at the first step); the root does. The broadcast
    # must still line up: the root announces its layout, the newcomer creates the slots, then the tensors travel
    for make in (lambda ps: torch.optim.SGD(ps, lr=0.1, momentum=0.9), lambda ps: torch.optim.Adam(ps, lr=1e-3)):
        m2 = torch.nn.Linear(5, 3)
        o2 = make(m2.parameters())
        if r == 0:
            for _ in range(2):
                o2.zero_grad()
                m2(torch.randn(4, 5)).sum().backward()
                o2.step()
        initializer.BroadcastGlobalVariablesOp(m2, o2)
        keys = sorted(k for p in m2.parameters() for k in o2.state[p].keys())
        check(len(keys) > 0 and ops.consensus(torch.tensor([float(len(keys))])), f"newcomer created the root's optimizer slots: {keys}")
        st = torch.cat([o2.state[p][k].detach().reshape(-1).float() for p in m2.parameters() for k in sorted(o2.state[p].keys())
                        if isinstance(o2.state[p][k], torch.Tensor)])
        check(ops.consensus(st) and float(st.abs().sum()) > 0, "lazily created optimizer state equals the root's after the broadcast")
        o2.zero_grad()
        torch.manual_seed(7)
        m2(torch.randn(4, 5)).sum().backward()
        o2.step()  # same data everywhere: identical state => identical parameters after the step
        check(ops.consensus(torch.cat([p.detach().reshape(-1) for p in m2.parameters()])), "a step from the broadcast state agrees on every worker")
    with torch.no_grad():
        for p in m.parameters():
            p.add_(float(r))
    cb = initializer.BroadcastGlobalVariablesCallback(m)
    cb.on_batch_end(0)
    cb.on_batch_end(1)  # only the first batch broadcasts
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    check(ops.consensus(flat), "callback broadcast at the end of the first batch")
    t = [torch.full((3,), float(r)), torch.full((2, 2), float(r + 1))]
    initializer.broadcast_variables(t, root=n - 1)
    check(bool(torch.all(t[0] == float(n - 1))) and bool(torch.all(t[1] == float(n))), "broadcast_variables from the last rank")
    check(timing.max_over_ranks(float(r)) == float(n - 1), "max_over_ranks")
    s = timing.ClockSampler(0).start()
    time.sleep(0.05)
    out = s.stop()  # no nvidia-smi on a CPU box: an empty summary, not an error
    check(set(out) >= {"sm_mhz", "reasons", "samples"}, "clock sampler summary keys")
    rep = trace.report()
    check(("init.broadcast" in rep) == trace.ENABLED, "trace report lists the scope when tracing is on")
    print(json.dumps({"ok": True, "traced": trace.ENABLED}), flush=True)


def cmd_compat(args):
    """A program written against the reference's module paths (`kungfu.torch.ops.clib` tables, `kungfu.tensorflow.*` names,
    `kungfu.utils.ema`, the ctypes loader) runs unchanged on torch tensors."""
    import kungfu  # noqa: F401
    import kungfu.torch as kft
    from kungfu.python import current_cluster_size, current_rank
    from kungfu.tensorflow.initializer import BroadcastGlobalVariablesOp  # noqa: F401
    from kungfu.tensorflow.ops import all_reduce as tf_all_reduce
    from kungfu.tensorflow.optimizers import SynchronousSGDOptimizer
    from kungfu.tensorflow.policy import BasePolicy, PolicyHook  # noqa: F401
    from kungfu.torch.ops.clib import all_gather_op_map, all_reduce_async_op_map, all_reduce_op_map, broadcast_async_op_map, ops as clib_ops
    from kungfu.torch.ops.collective import inplace_broadcast_async_op, wait_all_handles
    from kungfu.utils.ema import EMA

    n, r = current_cluster_size(), current_rank()
    check(kft.current_cluster_size() == n and SynchronousSGDOptimizer is kf.optimizers.SynchronousSGDOptimizer, "aliases are the same objects")
    x = torch.full((1000,), float(r + 1))
    y = torch.empty_like(x)
    all_reduce_op_map[x.type()](x, y, x.type(), "sum")  # the reference's call shape: (input, output, type, op)
    check(y[0].item() == n * (n + 1) / 2 and x[0].item() == r + 1, "table all-reduce, out of place")
    xi = torch.full((7,), r, dtype=torch.int64)
    all_reduce_op_map[xi.type()](xi, xi, xi.type(), "max")
    check(xi[0].item() == n - 1, "table all-reduce, in place, int64 max")
    g = torch.empty((n, 5))
    all_gather_op_map[x.type()](torch.full((5,), float(r)), g, x.type())
    check([g[q, 0].item() for q in range(n)] == [float(q) for q in range(n)], "table all-gather")
    hs = [all_reduce_async_op_map[x.type()](t, t, t.type(), "sum", f"a{i}") for i, t in enumerate(torch.ones(6, 300).unbind(0))]
    clib_ops.wait_all_handles(hs)
    ts = [torch.full((50 + i,), float(r * 10 + i)) for i in range(5)]
    hs = [broadcast_async_op_map[t.type()](t, t, t.type(), f"b{i}") for i, t in enumerate(ts)]
    hs += [inplace_broadcast_async_op(t, f"c{i}") for i, t in enumerate(ts)]
    wait_all_handles(hs)
    check(all(t[0].item() == float(i) and t[-1].item() == float(i) for i, t in enumerate(ts)), "asynchronous broadcasts deliver rank 0's values")
    check(float(tf_all_reduce(torch.ones(3))[0]) == n, "kungfu.tensorflow.ops.all_reduce on a torch tensor")
    e = EMA(0.5, scale_cap=2.0)
    check(e.update(4.0) == 4.0 and e.update(100.0) == 6.0, "EMA clamps a jump to value * scale_cap before blending")
    e.reset()
    check(e.update(1.0) == 1.0, "EMA reset")
    kf.run_barrier()
    print(json.dumps({"ok": True, "compat": True, "rank": r, "np": n}), flush=True)


def cmd_mismatch(args):
    """Workers that disagree on a tensor's size must get an error quickly, not wait for each other forever."""
    rt = get_runtime()
    ops.all_reduce(torch.ones(10 + 2 * rt.rank), name="bad")
    print("unexpected success", flush=True)


def cmd_bad(args):
    rt = get_runtime()
    if rt.rank == 0:
        time.sleep(0.5)
        sys.exit(7)
    time.sleep(60)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--batch", type=int, default=60)
    ap.add_argument("--schedule", default="1:3,2:3,4:3,2:3,1:3")
    ap.add_argument("--max-step", type=int, default=15)
    ap.add_argument("--dir", default="/tmp")
    ap.add_argument("--n-epochs", type=int, default=3)
    ap.add_argument("--restart", type=int, default=0)
    ap.add_argument("extra", nargs="*")
    args = ap.parse_args()
    globals()["cmd_" + args.what](args)


if __name__ == "__main__":
    main()
