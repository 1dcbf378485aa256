"""Small tools and adapters: rrun command generation, ModelArts env parsing, test-util kill, examples smoke, C ABI tool."""
import os
import subprocess
import sys
import time

import pytest
import torch

from test_runtime_multiproc import ROOT, kungfu_run, launch, ports


def test_rrun_command_generation():
    from kungfu_b200.cmd import rrun

    cmds = rrun.build_commands(16, "10.0.0.1:8,10.0.0.2:8", ["python3", "train.py", "--x", "a b"], user="ubuntu", watch=True, strategy="RING")
    assert len(cmds) == 2 and cmds[0][:3] == ["ssh", "-o", "StrictHostKeyChecking=no"] and cmds[1][3] == "ubuntu@10.0.0.2"
    assert "-self 10.0.0.2 -np 16 -H 10.0.0.1:8,10.0.0.2:8 -w -strategy RING python3 train.py --x 'a b'" in cmds[1][4]
    d = rrun.distribute_commands("10.0.0.1:8,10.0.0.2:8", ["hostname"])
    assert [c[3] for c in d] == ["10.0.0.1", "10.0.0.2"] and d[0][4] == "hostname"
    assert rrun.main(["-np", "2", "-H", "127.0.0.1:2", "--dry-run", "echo", "hi"]) == 0


def test_modelarts_env():
    from kungfu_b200.platforms import modelarts

    env = {"DLS_TASK_NUMBER": "2", "DLS_TASK_INDEX": "1", "BATCH_CUSTOM0_HOSTS": "10.1.0.1:6666", "BATCH_CUSTOM1_HOSTS": "10.1.0.2:6666"}
    idx, runners, workers = modelarts.parse_env(env, slots=2)
    assert idx == 1 and runners == ["10.1.0.1:38080", "10.1.0.2:38080"]
    assert workers == ["10.1.0.1:10000", "10.1.0.1:10001", "10.1.0.2:10000", "10.1.0.2:10001"]
    assert modelarts.kungfu_run_args(env, slots=2) == ["-np", "4", "-H", "10.1.0.1:2,10.1.0.2:2", "-self", "10.1.0.2"]


def test_kill_worker_via_control_message():
    """≙ `kungfu-test-util -kill ip:port`: the `exit` control message stops a worker (exit 0); its siblings then fail and the
    launcher reports a non-zero status (fail-fast)."""
    rport, prange = ports()
    base = int(prange.split("-")[0])
    cmd = [kungfu_run(), "-q", "-np", "2", "-H", "127.0.0.1:2", "-port", str(rport), "-port-range", prange, sys.executable,
           os.path.join(ROOT, "tests", "cpu_worker.py"), "bad"]
    # use sleeping workers: rank 0 of `bad` exits 7 after 0.5 s, so instead run a sleeper via python -c
    code = "import sys,time; sys.path.insert(0,%r); import kungfu_b200.torch as kf; kf.run_barrier(); print('up', flush=True); time.sleep(60)" % ROOT
    cmd = cmd[:-2] + ["-c", code]
    p = subprocess.Popen(cmd, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    from kungfu_b200 import _kfrt as rt

    target = f"127.0.0.1:{base + 1}"
    for _ in range(300):
        if rt.ping(target) >= 0:
            break
        time.sleep(0.1)
    time.sleep(1.0)
    assert subprocess.call([sys.executable, "-m", "kungfu_b200.cmd.test_util", "--kill", target], cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT)) == 0
    t0 = time.time()
    while rt.ping(target) >= 0 and time.time() - t0 < 10:
        time.sleep(0.1)
    assert rt.ping(target) < 0, "worker should be gone"
    p.terminate()
    p.wait(timeout=20)


def test_fake_agent_c_abi():
    """C++ program over the C ABI under the launcher, ring strategy (≙ run-integration-tests.sh)."""
    from kungfu_b200.cmd import _bin

    r = launch(3, "unused", expect_rc=None, flags=())  # warm: ensures binaries are built
    rport, prange = ports()
    out = subprocess.run([kungfu_run(), "-q", "-np", "3", "-H", "127.0.0.1:3", "-port", str(rport), "-port-range", prange, "-strategy", "RING",
                          _bin("kungfu-fake-agent"), "test"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert out.returncode == 0 and "fake-agent ok np=3" in out.stdout, out.stdout[-2000:]
    out = subprocess.run([kungfu_run(), "-q", "-np", "2", "-H", "127.0.0.1:2", "-port", str(rport), "-port-range", prange,
                          _bin("kungfu-fake-agent"), "bench", "slp"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert out.returncode == 0 and "RESULT:" in out.stdout, out.stdout[-2000:]
    # fake trainer, asynchronous flavour: all 161 ResNet-50 tensors in flight through the callback form of the C ABI
    rport, prange = ports()
    out = subprocess.run([kungfu_run(), "-q", "-np", "3", "-H", "127.0.0.1:3", "-port", str(rport), "-port-range", prange,
                          _bin("kungfu-fake-agent"), "bench-async", "resnet50"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert out.returncode == 0 and "RESULT:" in out.stdout and "Img/sec" in out.stdout, out.stdout[-2000:]


def test_fake_agent_survives_a_lost_worker_through_the_c_abi():
    """`kungfu-fake-agent shrink 1@5` under `kungfu-run -w -shrink-on-failure`: a C program on the C ABI (kungfu_all_reduce fails ->
    kungfu_recover -> repeat the step) loses rank 1 inside step 5 and finishes its 20 steps with two workers."""
    from kungfu_b200.cmd import kungfu_run_path

    agent = os.path.join(os.path.dirname(kungfu_run_path()), "kungfu-fake-agent")
    rport, prange = ports()
    cmd = [kungfu_run(), "-q", "-w", "-shrink-on-failure", "-np", "3", "-H", "127.0.0.1:4", "-port", str(rport), "-port-range", prange,
           "-builtin-config-port", str(9200 + os.getpid() % 90), agent, "shrink", "1@5"]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, KUNGFU_CONFIG_LOG_LEVEL="WARN"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("fake-agent shrink ok") == 2 and r.stdout.count("recovered=1 sizes=2..3 final=2") == 2, r.stdout[-3000:]


def test_examples_and_microbench_run():
    rport, prange = ports()
    env = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN")
    base = [kungfu_run(), "-q", "-np", "2", "-H", "127.0.0.1:2", "-port", str(rport), "-port-range", prange, sys.executable]
    r = subprocess.run(base + [os.path.join(ROOT, "examples", "torch_mnist_example.py"), "--epochs", "1"], env=env, cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
    assert r.returncode == 0 and "acc" in r.stdout, r.stdout[-2000:]
    r = subprocess.run(base + ["-m", "kungfu_b200.benchmarks", "--model", "slp", "--steps", "3"], env=env, cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
    assert r.returncode == 0 and "RESULT:" in r.stdout, r.stdout[-2000:]
    r = subprocess.run(base + ["-m", "kungfu_b200.examples"], env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=180)  # self-checking: all-reduce of ones, closed-form S-SGD / SMA, PairAveraging envelope
    assert r.returncode == 0 and "examples ok" in r.stdout, r.stdout[-2000:]
    r = subprocess.run(base + [os.path.join(ROOT, "examples", "adapt_strategy.py")], env=env, cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
    assert r.returncode == 0, r.stdout[-2000:]


def test_checkpoint_roundtrip(tmp_path):
    from kungfu_b200 import checkpoint

    m = torch.nn.Linear(3, 2)
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)
    m(torch.ones(1, 3)).sum().backward()
    opt.step()
    path = str(tmp_path / "c.pt")
    assert checkpoint.save(path, m, opt, step=7, epoch=2, trained_samples=99, extra={"k": 1}) == path
    m2 = torch.nn.Linear(3, 2)
    opt2 = torch.optim.SGD(m2.parameters(), lr=0.1, momentum=0.9)
    meta = checkpoint.load(path, m2, opt2)
    assert meta["found"] and meta["step"] == 7 and meta["epoch"] == 2 and meta["trained_samples"] == 99 and meta["extra"] == {"k": 1}
    assert all(torch.equal(a, b) for a, b in zip(m.parameters(), m2.parameters()))
    assert checkpoint.load(str(tmp_path / "missing.pt"), m2)["found"] is False


def test_policy_hook_scheduled_batch_size():
    """≙ tests/python/unit/test_tensorflow_policy.py with scheduled_batch_size_policy.py."""
    from kungfu_b200.elastic import PolicyHook, ScheduledBatchSizePolicy

    hook = PolicyHook([ScheduledBatchSizePolicy({0: 10, 1: 20, 2: 50})], epoch_size=100, epoch_num=3, init_batch_size=10)
    hook.begin()
    sizes = []
    while hook.before_step():
        sizes.append(hook.batch_size)
        hook.after_step()
    hook.end()
    assert sizes == [10] * 10 + [20] * 5 + [50] * 2
    assert hook.vars["kungfu_trained_samples"] == 300 and hook.vars["epoch"] == 3


def test_elastic_data_adaptor():
    from kungfu_b200.elastic import ElasticDataAdaptor

    data, labels = torch.arange(100).float().view(100, 1), torch.arange(100)
    ad = ElasticDataAdaptor(data, labels, batch_size=8)
    x, y = ad.next_batch()
    assert y.tolist() == list(range(8)) and ad.offset == 8
    ad.set_offset(96)
    _, y = ad.next_batch()
    assert y.tolist() == [96, 97, 98, 99, 0, 1, 2, 3]


def test_monitoring_benchmark_reports_overhead_and_values():
    """benchmarks/monitoring/benchmark.py: plain S-SGD and both monitors in one job, one RESULT line per arm and interval, live monitor values
    (workers see different data, so variance and noise scale are not zero), JSON summary."""
    import json

    rport, prange = ports()
    cmd = [kungfu_run(), "-q", "-np", "2", "-H", "127.0.0.1:2", "-port", str(rport), "-port-range", prange, sys.executable,
           os.path.join(ROOT, "benchmarks", "monitoring", "benchmark.py"), "--no-cuda", "--model", "slp", "--num-iters", "2",
           "--num-batches-per-iter", "4", "--num-warmup-batches", "2", "--intervals", "1,3"]
    e = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN")
    r = subprocess.run(cmd, cwd=ROOT, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("RESULT:") == 5, r.stdout[-3000:]  # sync-sgd + 2 monitors x 2 intervals
    line = [l for l in r.stdout.splitlines() if '"bench": "monitoring"' in l][-1]
    summary = json.loads(line[line.index("{"):])
    rows = summary["rows"]
    assert [r_["arm"] for r_ in rows] == ["sync-sgd", "variance", "variance", "noise-scale", "noise-scale"] and summary["np"] == 2
    assert all(r_["img_per_sec"] > 0 for r_ in rows) and rows[1]["value"] > 0 and rows[3]["value"] != 0


# ---- the reference's module paths and foreign-binding entry points ------------------------------------------------------------------
def test_reference_module_paths_resolve_to_one_implementation():
    """`kungfu.X[.Y]` maps onto `kungfu_b200` by a meta-path finder: the SAME module objects (no second copy of a class), including the
    `kungfu.tensorflow.*` paths, submodules the alias list never named, and `python -m` of aliased packages."""
    import importlib

    import kungfu
    import kungfu_b200

    assert kungfu.__version__ == kungfu_b200.__version__
    pairs = {"kungfu.torch.optimizers.sync_sgd": "kungfu_b200.torch.optimizers.sync_sgd", "kungfu.torch.ops.collective": "kungfu_b200.torch.ops.collective",
             "kungfu.torch.ops.clib": "kungfu_b200.torch.ops.clib", "kungfu.utils.ema": "kungfu_b200.utils.ema", "kungfu.loader": "kungfu_b200.loader",
             "kungfu._utils": "kungfu_b200.utils", "kungfu.tensorflow.ops": "kungfu_b200.ops", "kungfu.tensorflow.ops.adapt": "kungfu_b200.ops.adapt",
             "kungfu.tensorflow.optimizers.async_sgd": "kungfu_b200.torch.optimizers.async_sgd", "kungfu.tensorflow.optimizers.keras": "kungfu_b200.torch.optimizers",
             "kungfu.tensorflow.initializer.keras": "kungfu_b200.initializer", "kungfu.tensorflow.policy.policy_hook": "kungfu_b200.elastic.policy",
             "kungfu.tensorflow.hooks.elastic": "kungfu_b200.elastic.hooks", "kungfu.tensorflow.variables": "kungfu_b200.variables",
             "kungfu.tensorflow.v1.helpers.mnist": "kungfu_b200.helpers.mnist", "kungfu.tensorflow.v1.datasets.adaptor": "kungfu_b200.elastic.datasets",
             "kungfu.tensorflow.v1.benchmarks.model_sizes": "kungfu_b200.benchmarks.model_sizes", "kungfu.parallel.engine": "kungfu_b200.parallel.engine"}
    for alias, real in pairs.items():
        assert importlib.import_module(alias) is importlib.import_module(real), alias
    with pytest.raises(ModuleNotFoundError):
        importlib.import_module("kungfu.no_such_module")
    env = dict(os.environ, PYTHONPATH=ROOT)
    for mod, want in (("kungfu.info", '"native_runtime": true'), ("kungfu.tensorflow.v1.examples", "examples ok")):
        r = subprocess.run([sys.executable, "-m", mod], env=env, cwd="/tmp", stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
        assert r.returncode == 0 and want in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("np", [1, 3])
def test_program_written_against_reference_paths(np):
    from test_runtime_multiproc import json_lines, launch

    out = launch(np, "compat").stdout
    assert len([j for j in json_lines(out) if j.get("compat")]) == np, out[-3000:]


@pytest.mark.parametrize("np", [1, 2, 4])
def test_c_abi_through_ctypes_without_torch(np):
    """kungfu_b200/loader.py: the C ABI loaded with ctypes by a worker that never imports torch (sync and callback all-reduce, int max,
    p2p save / request incl. the negative case)."""
    from test_runtime_multiproc import json_lines, kungfu_run, ports

    rport, prange = ports()
    cmd = [kungfu_run(), "-q", "-np", str(np), "-H", f"127.0.0.1:{np}", "-port", str(rport), "-port-range", prange, sys.executable,
           os.path.join(ROOT, "tests", "capi_ctypes_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, KUNGFU_CONFIG_LOG_LEVEL="WARN"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-3000:]
    assert len([j for j in json_lines(r.stdout) if j.get("ctypes") and j.get("uid")]) == np


def test_c_program_links_against_libkungfu_comm(tmp_path):
    """examples/c/allreduce.c compiled as C99 against capi.h and lib/libkungfu-comm.so (no Python in the link), run under the launcher."""
    from kungfu_b200 import _build

    _build.build_runtime()
    lib = _build.PKG / "lib"
    assert (lib / "libkungfu-comm.so").exists()
    exe = tmp_path / "c_allreduce"
    subprocess.run(["gcc", "-std=c99", "-D_DEFAULT_SOURCE", "-Wall", "-Werror", "-I", str(_build.CSRC / "runtime"), os.path.join(ROOT, "examples", "c", "allreduce.c"),
                    "-L", str(lib), "-lkungfu-comm", f"-Wl,-rpath,{lib}", "-o", str(exe)], check=True)
    deps = subprocess.run(["ldd", str(exe)], stdout=subprocess.PIPE, text=True).stdout
    assert "libkungfu-comm.so" in deps and "python" not in deps.lower()
    rport, prange = ports()
    r = subprocess.run([kungfu_run(), "-q", "-np", "4", "-H", "127.0.0.1:4", "-port", str(rport), "-port-range", prange, "-strategy", "RING", str(exe)],
                       env=dict(os.environ, KUNGFU_CONFIG_LOG_LEVEL="WARN"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and "c worker ok np=4" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("np,strategy", [(1, "AUTO"), (3, "STAR"), (4, "BINARY_TREE_STAR")])
def test_cpp_application_on_the_public_header(tmp_path, np, strategy):
    """examples/cpp/adapt_tree.cpp on `kungfu.hpp` (header-only `kungfu::Peer` over the C ABI): latencies -> AllGatherTransform with a
    minimum spanning tree -> consensus -> SetTree -> synchronous and callback all-reduces, p2p with a missing blob."""
    from kungfu_b200 import _build

    _build.build_runtime()
    lib = _build.PKG / "lib"
    exe = tmp_path / "adapt_tree"
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", str(_build.CSRC / "runtime"), os.path.join(ROOT, "examples", "cpp", "adapt_tree.cpp"),
                    "-L", str(lib), "-lkungfu-comm", f"-Wl,-rpath,{lib}", "-pthread", "-o", str(exe)], check=True)
    rport, prange = ports()
    r = subprocess.run([kungfu_run(), "-q", "-np", str(np), "-H", f"127.0.0.1:{np}", "-port", str(rport), "-port-range", prange, "-strategy", strategy, str(exe)],
                       env=dict(os.environ, KUNGFU_CONFIG_LOG_LEVEL="WARN"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and f"adapt_tree ok np={np} tree=[0" in r.stdout, r.stdout[-3000:]


def test_minimum_spanning_tree_of_the_public_header(tmp_path):
    """`kungfu::minimum_spanning_tree` against brute force on random dense graphs (reference: tests/cpp/unit/test_mst.cpp, n = 1..10)."""
    src = tmp_path / "mst.cpp"
    src.write_text(r'''
#include <algorithm>
#include <cstdio>
#include <random>
#include <vector>
#include "kungfu.hpp"
static double tree_cost(int n, const std::vector<double>& w, const std::vector<int32_t>& f) {
    double c = 0;
    for (int i = 0; i < n; ++i) if (f[i] != i) c += (w[i * n + f[i]] + w[f[i] * n + i]) / 2;
    return c;
}
static bool is_tree(int n, const std::vector<int32_t>& f, int root) {
    if (f[root] != root) return false;
    for (int i = 0; i < n; ++i) { int v = i, hops = 0; while (v != root && hops++ <= n) v = f[v]; if (v != root) return false; }
    return true;
}
int main() {
    std::mt19937 rng(7);
    for (int n = 1; n <= 7; ++n)
        for (int rep = 0; rep < 20; ++rep) {
            std::vector<double> w(n * n);
            for (auto& x : w) x = 1 + rng() % 1000;
            const int root = rng() % n;
            auto f = kungfu::minimum_spanning_tree(n, w.data(), root);
            if (!is_tree(n, f, root)) { std::printf("not a tree n=%d\n", n); return 1; }
            // brute force: every father array that is a tree rooted at root
            double best = 1e18;
            std::vector<int32_t> g(n, 0);
            long total = 1;
            for (int i = 0; i < n; ++i) total *= n;
            for (long code = 0; code < total; ++code) {
                long c = code;
                for (int i = 0; i < n; ++i) { g[i] = c % n; c /= n; }
                if (is_tree(n, g, root)) best = std::min(best, tree_cost(n, w, g));
            }
            if (tree_cost(n, w, f) > best + 1e-9) { std::printf("n=%d cost %f > optimum %f\n", n, tree_cost(n, w, f), best); return 1; }
        }
    std::printf("mst ok\n");
    return 0;
}
''')
    from kungfu_b200 import _build

    exe = tmp_path / "mst"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(_build.CSRC / "runtime"), str(src), "-L", str(_build.PKG / "lib"), "-lkungfu-comm",
                    f"-Wl,-rpath,{_build.PKG / 'lib'}", "-pthread", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0 and "mst ok" in r.stdout, r.stdout
