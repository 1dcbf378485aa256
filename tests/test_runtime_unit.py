"""Unit tests of the native runtime pieces (no processes, no GPU).

Mirrors the reference's Go unit tests (srcs/go/plan/{topology,cluster,hostspec}_test.go, plan/hostfile/hostfile_test.go,
plan/graph/graph_test.go, store/versionedstore_test.go) and C++ gtest units (tests/cpp/unit/test_{operations,mst,
placement}.cpp) plus the python unit tests (tests/python/unit/test_op.py: counter, EMA trace, step_based_schedule).
"""
import itertools
import json
import random

import pytest
import torch

from kungfu_b200 import _kfrt as rt
from kungfu_b200 import ops
from kungfu_b200.utils import dtypes


def peers(spec):
    """spec: list of (host index, slots) -> peer list string"""
    out = []
    for h, slots in spec:
        out += [f"10.0.0.{h + 1}:{10000 + i}" for i in range(slots)]
    return ",".join(out)


# ---- plan --------------------------------------------------------------------------------------------
def test_gen_peer_list_and_runners():
    pl = rt.gen_peer_list("192.168.1.1:2,192.168.1.2:3", 4, "10000-11000")
    assert pl == ["192.168.1.1:10000", "192.168.1.1:10001", "192.168.1.2:10000", "192.168.1.2:10001"]
    assert rt.gen_runner_list("192.168.1.1:2,192.168.1.2:3", 38080) == ["192.168.1.1:38080", "192.168.1.2:38080"]
    assert rt.host_list_cap("127.0.0.1:4,127.0.0.2:4") == 8
    with pytest.raises(RuntimeError):
        rt.gen_peer_list("127.0.0.1:2", 3, "10000-11000")  # no enough capacity
    with pytest.raises(RuntimeError):
        rt.gen_peer_list("127.0.0.1:4", 4, "10000-10001")  # port range too small


def test_peer_list_info():
    pl = "10.0.0.1:10000,10.0.0.1:10001,10.0.0.2:10000,10.0.0.2:10001,10.0.0.2:10002"
    info = rt.peer_list_info(pl, "10.0.0.2:10001")
    assert info["rank"] == 3 and info["local_rank"] == 1 and info["local_size"] == 3 and info["host_count"] == 2
    assert info["masters"] == [0, 2] and info["parents"] == [0, 0, 2, 2, 2]


def test_hostfile_openmpi_syntax():
    text = "# comment\n10.0.0.1 slots=4\n10.0.0.2 slots=2 max_slots=8\n\n10.0.0.3\n"
    hl = rt.parse_hostfile(text)
    assert hl.split(",")[0].startswith("10.0.0.1:4") and hl.split(",")[1].startswith("10.0.0.2:2") and hl.split(",")[2].startswith("10.0.0.3:1")


def test_cluster_json_validate_resize():
    c = rt.cluster_json("10.0.0.1:38080,10.0.0.2:38080", "10.0.0.1:10000,10.0.0.2:10000")
    d = json.loads(c)
    assert set(d) == {"Runners", "Workers"} and d["Workers"][0]["Port"] == 10000 and isinstance(d["Workers"][0]["IPv4"], int)
    rt.cluster_validate(c)
    # growth goes to the least loaded runner host, with port = max + 1; shrink keeps the prefix
    grown = rt.cluster_from_json(rt.cluster_resize(c, 5))["workers"]
    assert grown[:2] == ["10.0.0.1:10000", "10.0.0.2:10000"] and len(grown) == 5
    assert sorted(grown[2:]) == ["10.0.0.1:10001", "10.0.0.1:10002", "10.0.0.2:10001"]
    shrunk = rt.cluster_from_json(rt.cluster_resize(rt.cluster_resize(c, 5), 1))["workers"]
    assert shrunk == ["10.0.0.1:10000"]
    with pytest.raises(RuntimeError):  # worker without a runner on its host
        rt.cluster_validate(rt.cluster_json("10.0.0.1:38080", "10.0.0.9:10000"))
    with pytest.raises(RuntimeError):  # duplicated worker
        rt.cluster_validate(rt.cluster_json("10.0.0.1:38080", "10.0.0.1:10000,10.0.0.1:10000"))
    ver, cj = rt.stage_from_json(rt.stage_json(7, c))
    assert ver == 7 and json.loads(cj) == d


def test_even_partition():
    for n, k in [(10, 3), (7, 7), (5, 8), (1 << 20, 5)]:
        parts = [rt.even_partition(n, k, i) for i in range(k)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(k - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


# ---- graphs ------------------------------------------------------------------------------------------
TOPOS = [[(0, 1)], [(0, 4)], [(0, 2), (1, 2)], [(0, 3), (1, 1), (2, 4)], [(0, 1), (1, 1), (2, 1), (3, 1), (4, 1)]]


@pytest.mark.parametrize("strategy", ["STAR", "MULTI_STAR", "RING", "CLIQUE", "TREE", "BINARY_TREE", "BINARY_TREE_STAR", "MULTI_BINARY_TREE_STAR", "AUTO"])
@pytest.mark.parametrize("topo", TOPOS)
def test_strategy_graphs_are_rooted_trees(strategy, topo):
    """Every generator must yield, per strategy, a broadcast graph that is a tree reaching all peers and a reduce graph
    that is its reverse with self loops (≙ plan/topology_test.go:14-98)."""
    pl = peers(topo)
    n = len(pl.split(","))
    pairs = rt.gen_strategy_graphs(pl, strategy, "global")
    assert len(pairs) >= 1
    for red, bc in pairs:
        assert red["n"] == bc["n"] == n
        roots = [i for i in range(n) if all(e[1] != i for e in bc["edges"])]
        assert len(roots) == 1
        assert rt.is_rooted_tree(n, bc["edges"], roots[0])
        if strategy != "RING":
            assert sorted((b, a) for a, b in bc["edges"]) == sorted(map(tuple, red["edges"]))
        assert red["self_loops"] == list(range(n))
    if strategy == "RING" or strategy == "CLIQUE":
        assert len(pairs) == n
    loc = rt.gen_strategy_graphs(pl, "AUTO", "local")
    assert len(loc) == 1 and len(loc[0][1]["edges"]) == n - len(topo)  # a star per host


def test_forest_array():
    ok, roots, g = rt.forest_to_graph([0, 0, 1, 1, 0])
    assert ok and roots == 1 and sorted(map(tuple, g["edges"])) == [(0, 1), (0, 4), (1, 2), (1, 3)]
    ok, roots, _ = rt.forest_to_graph([1, 0, 2])  # cycle 0 <-> 1
    assert not ok
    ok, roots, _ = rt.forest_to_graph([0, 1, 2])
    assert ok and roots == 3


def test_mst_random_permutation_graphs():
    """≙ tests/cpp/unit/test_mst.cpp: hide a cheap Hamiltonian path in a random permutation; the MST must find it."""
    rng = random.Random(7)
    for n in range(1, 11):
        perm = list(range(n))
        rng.shuffle(perm)
        w = [[100.0 + rng.random() for _ in range(n)] for _ in range(n)]
        for a, b in zip(perm, perm[1:]):
            w[a][b] = w[b][a] = 1.0
        edges = rt.minimum_spanning_tree(list(itertools.chain(*w)), n)
        assert len(edges) == n - 1
        assert sorted(tuple(sorted(e)) for e in edges) == sorted(tuple(sorted(p)) for p in zip(perm, perm[1:]))
        father = rt.mst_forest_array(list(itertools.chain(*w)), n, 0)
        ok, roots, _ = rt.forest_to_graph(father)
        assert ok and roots == 1 and father[0] == 0


# ---- reduce kernels ----------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int32, torch.int64, torch.uint8, torch.int8, torch.int16])
@pytest.mark.parametrize("op", ["sum", "min", "max", "prod"])
def test_transform2_matches_torch(dtype, op):
    g = torch.Generator().manual_seed(1)
    n = 1037
    if dtype.is_floating_point:
        x, y = (torch.randn(n, generator=g) * 3).to(dtype), (torch.randn(n, generator=g) * 3).to(dtype)
    else:
        x, y = torch.randint(0, 11, (n,), generator=g).to(dtype), torch.randint(0, 11, (n,), generator=g).to(dtype)
    out = torch.empty_like(x)
    rt.transform2(out.data_ptr(), x.data_ptr(), y.data_ptr(), n, dtypes.dtype_code(dtype), dtypes.op_code(op)[0])
    if dtype in (torch.float16, torch.bfloat16):
        xf, yf = x.float(), y.float()
        ref = {"sum": xf + yf, "min": torch.minimum(xf, yf), "max": torch.maximum(xf, yf), "prod": xf * yf}[op].to(dtype)
    else:
        ref = {"sum": x + y, "min": torch.minimum(x, y), "max": torch.maximum(x, y), "prod": x * y}[op]
    assert torch.equal(out, ref)
    # in place on the first operand (how the session accumulates)
    acc = x.clone()
    rt.transform2(acc.data_ptr(), acc.data_ptr(), y.data_ptr(), n, dtypes.dtype_code(dtype), dtypes.op_code(op)[0])
    assert torch.equal(acc, ref)


def test_half_conversions_roundtrip():
    vals = torch.tensor([0.0, -0.0, 1.0, -2.5, 65504.0, 1e-7, 6.1e-5, 3.14159, float("inf"), -float("inf")])
    for v in vals.tolist():
        h = torch.tensor([v]).half()
        assert rt.float_to_half(v) == int(h.view(torch.int16).item()) & 0xFFFF
        assert rt.half_to_float(int(h.view(torch.int16).item()) & 0xFFFF) == float(h.float())
        b = torch.tensor([v]).bfloat16()
        assert rt.float_to_bf16(v) == int(b.view(torch.int16).item()) & 0xFFFF
        assert rt.bf16_to_float(int(b.view(torch.int16).item()) & 0xFFFF) == float(b.float())


# ---- numa placement ----------------------------------------------------------------------------------
def test_select_cpus():
    numa = [list(range(0, 8)), list(range(8, 16))]
    got = [rt.select_cpus(numa, r, 4) for r in range(4)]
    assert got[0] == [0, 1, 2, 3] and got[1] == [4, 5, 6, 7] and got[2] == [8, 9, 10, 11] and got[3] == [12, 13, 14, 15]
    assert rt.select_cpus(numa, 0, 1) == list(range(0, 8))
    assert rt.select_cpus([], 0, 1) == [] and rt.select_cpus(numa, 5, 4) == []
    assert len(rt.numa_topology()) >= 1


# ---- small python ops --------------------------------------------------------------------------------
def test_counter_ema_schedule_fuse():
    c = ops.counter()
    assert [c() for _ in range(3)] == [0, 1, 2]
    c = ops.counter(5, 2)
    assert [c() for _ in range(3)] == [5, 7, 9]
    ema = ops.exponential_moving_average(0.9)
    trace = [ema.update(x) for x in (1.0, 2.0, 3.0)]
    assert trace[0] == 1.0 and abs(trace[1] - 1.1) < 1e-12 and abs(trace[2] - (0.9 * 1.1 + 0.3)) < 1e-12
    sch = ops.StepBasedSchedule("1:3,2:3,4:3")
    assert [sch(i) for i in range(10)] == [1, 1, 1, 2, 2, 2, 4, 4, 4, 1]
    with pytest.raises(ValueError):
        ops.StepBasedSchedule("1:3", strict=True)(5)
    with pytest.raises(ValueError):
        ops.StepBasedSchedule("bad")
    ts = [torch.arange(6.0).view(2, 3), torch.ones(4), torch.zeros(1, 1)]
    f = ops.fuse(ts)
    back = ops.defuse(f, [t.shape for t in ts])
    assert f.numel() == 11 and all(torch.equal(a, b) for a, b in zip(ts, back))
    ns = ops.NoiseScale(0.5)
    v = ns.update(32, 128, 10.0, 4.0)
    G = (128 * 4.0 - 32 * 10.0) / 96
    S = (10.0 - 4.0) / (1 / 32 - 1 / 128)
    assert abs(v - S / G) < 1e-9


def test_single_process_mode():
    """No launcher env: rank 0 of 1, collectives are identities (≙ tests/python/unit/test_op.py rank()==0 etc.)."""
    import kungfu_b200.python as kfp
    import kungfu_b200.torch as kf

    assert kfp.current_rank() == 0 and kfp.current_cluster_size() == 1 and not kfp.detached()
    x = torch.arange(5.0)
    assert torch.equal(kf.ops.all_reduce_fn(x), x)
    assert kf.ops.all_gather(x).shape == (1, 5)
    kfp.run_barrier()
    assert kfp.resize() == (False, False)
    p = torch.nn.Parameter(torch.ones(3))
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD([p], lr=0.5), named_parameters=[("p", p)])
    assert isinstance(opt, torch.optim.SGD)
    p.grad = torch.ones(3)
    opt.step()
    assert torch.equal(p.data, torch.full((3,), 0.5))
    sched = torch.optim.lr_scheduler.StepLR(opt, 1, 0.1)  # LR schedulers accept the wrapped optimizer
    opt.step()
    sched.step()
    assert abs(opt.param_groups[0]["lr"] - 0.05) < 1e-12


def test_flat_layout_buckets():
    from kungfu_b200.parallel.flat import FlatLayout

    ps = [(f"p{i}", torch.nn.Parameter(torch.zeros(n))) for i, n in enumerate([10, 3000, 70000, 5, 123456, 64])]
    lay = FlatLayout(ps, torch.bfloat16, 100 * 1024)
    assert [e.name for e in lay.entries] == [f"p{i}" for i in reversed(range(6))]  # reverse registration order
    for e in lay.entries:
        assert e.offset % 64 == 0  # 128-byte alignment in bf16 elements
    assert sum(len(b.entries) for b in lay.buckets) == 6
    for a, b in zip(lay.buckets, lay.buckets[1:]):
        assert a.start + a.numel == b.start
    assert lay.buckets[-1].start + lay.buckets[-1].numel == lay.total
    for b in lay.buckets:
        assert (b.numel * 2) % 16 == 0
        lo0, hi0 = lay.shard_range(b, 8, 0)
        lo7, hi7 = lay.shard_range(b, 8, 7)
        assert lo0 == b.start and hi7 <= b.start + b.numel and hi0 - lo0 <= lay.shard_elems(b, 8)
    assert lay.digest() == FlatLayout(ps, torch.bfloat16, 100 * 1024).digest()


def test_s2d_stem_equals_plain_conv():
    """The space-to-depth stem is the same function (and has the same weight gradient) as the 7x7/2 convolution."""
    import torch.nn.functional as F

    from kungfu_b200.models.resnet import stem_conv_s2d

    torch.manual_seed(0)
    x = torch.randn(2, 3, 32, 40).contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, 3, 7, 7, requires_grad=True)
    y0, y1 = F.conv2d(x, w, None, 2, 3), stem_conv_s2d(x, w)
    assert y0.shape == y1.shape and torch.allclose(y0, y1, atol=1e-4)
    g = torch.randn_like(y0)
    (g0,) = torch.autograd.grad(y0, w, g, retain_graph=True)
    (g1,) = torch.autograd.grad(y1, w, g)
    assert torch.allclose(g0, g1, atol=1e-3)


def test_cpp_unit_test_binary():
    """The native runtime has its own C++ unit tests (csrc/tests/kungfu_unit_tests.cpp: reducers, plan, topology generators,
    MST, NUMA placement, stores, single-process Peer) — the counterpart of the reference's gtest suite; run the binary."""
    import subprocess

    from kungfu_b200 import _build

    _build.build_runtime()
    exe = _build.BIN / "kungfu-unit-tests"
    assert exe.exists()
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert "0 failures" in r.stdout
