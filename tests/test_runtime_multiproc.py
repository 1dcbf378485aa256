"""Multi-process integration tests on loopback under our own launcher (no GPU).

Reference model: `kungfu-run -np N -H 127.0.0.1:N` spawns N workers with distinct ports; every
integration test is an ordinary program run under the launcher for np in {1..4}, C++ integration × all strategies
(scripts/tests/run-integration-tests.sh:30-38).
"""
import itertools
import json
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "cpu_worker.py")
_counter = itertools.count()


def kungfu_run():
    from kungfu_b200.cmd import kungfu_run_path

    return kungfu_run_path()


def ports():
    """Disjoint (runner port, worker port range) per launch."""
    k = next(_counter)
    # below the kernel's ephemeral range (32768+): a listening port up there can collide with the source port of any outgoing
    # connection on the machine ("Address already in use" once in a few hundred launches)
    base = 11000 + ((os.getpid() * 37 + k * 64) % 18000)
    return base, f"{base + 1}-{base + 60}"


def launch(np, what, extra=(), flags=(), timeout=120, slots=None, env=None, expect_rc=0):
    rport, prange = ports()
    cmd = [kungfu_run(), "-q", "-np", str(np), "-H", f"127.0.0.1:{slots or max(np, 1)}", "-port", str(rport), "-port-range", prange,
           *flags, sys.executable, WORKER, what, *extra]
    e = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN")
    e.update(env or {})
    r = subprocess.run(cmd, cwd=ROOT, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    if expect_rc is not None:
        assert r.returncode == expect_rc, f"rc={r.returncode}\n{r.stdout[-4000:]}"
    return r


def json_lines(out):
    res = []
    for line in out.splitlines():
        i = line.find("{")
        if i >= 0:
            try:
                res.append(json.loads(line[i:]))
            except ValueError:
                pass
    return res


@pytest.mark.parametrize("np", [1, 2, 3, 4])
def test_collectives_default_strategy(np):
    r = launch(np, "collectives")
    assert len([j for j in json_lines(r.stdout) if j.get("ok")]) == np, r.stdout[-3000:]


@pytest.mark.parametrize("strategy", ["STAR", "MULTI_STAR", "RING", "CLIQUE", "TREE", "BINARY_TREE", "BINARY_TREE_STAR", "MULTI_BINARY_TREE_STAR", "AUTO"])
def test_collectives_all_strategies(strategy):
    r = launch(4, "collectives", flags=("-strategy", strategy))
    assert len([j for j in json_lines(r.stdout) if j.get("ok")]) == 4, r.stdout[-3000:]


def test_collectives_name_hash_and_small_chunks():
    """Strategy chosen by tensor-name hash + 64 KiB chunks (many concurrent chunk tasks on the elastic pool)."""
    r = launch(3, "collectives", flags=("-strategy", "RING"),
               env={"KUNGFU_CONFIG_STRATEGY_HASH_METHOD": "NAME", "KUNGFU_CONFIG_CHUNK_BYTES": "65536"})
    assert len([j for j in json_lines(r.stdout) if j.get("ok")]) == 3


def test_tcp_only_transport():
    r = launch(3, "collectives", env={"KUNGFU_CONFIG_USE_UNIX_SOCK": "false"})
    assert len([j for j in json_lines(r.stdout) if j.get("ok")]) == 3


@pytest.mark.parametrize("np", [1, 2, 4])
def test_p2p_store(np):
    r = launch(np, "p2p")
    assert len([j for j in json_lines(r.stdout) if j.get("ok")]) == np, r.stdout[-3000:]


def test_many_chunks_in_flight_do_not_starve_the_pool():
    """ResNet-50 sized messages (up to nine 1 MiB chunks per tensor, 161 tensors) on the default tree strategy: every chunk
    task blocks on its peer's task, so each needs its own worker thread. Regression test for a lost-wakeup style
    under-provisioning in `Pool::submit` that deadlocked roughly one run in three."""
    for _ in range(3):
        rport, prange = ports()
        cmd = [kungfu_run(), "-q", "-np", "4", "-H", "127.0.0.1:4", "-port", str(rport), "-port-range", prange, "-strategy", "BINARY_TREE_STAR",
               sys.executable, "-m", "kungfu_b200.benchmarks", "--method", "CPU", "--warmup-steps", "1", "--steps", "2"]
        r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN"), stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=90)
        assert r.returncode == 0 and "RESULT:" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("np,strategy", [(4, "BINARY_TREE_STAR"), (3, "RING"), pytest.param(4, "MULTI_BINARY_TREE_STAR", marks=pytest.mark.extended)])
def test_randomised_soak(np, strategy):
    """150 pseudo-random operations (all-reduce over dtypes / ops / sizes up to several chunks, broadcast from random roots,
    all-gather, threaded concurrent collectives, P2P requests, consensus, subset all-reduce), every result checked."""
    r = launch(np, "stress", extra=["150"], flags=["-strategy", strategy], timeout=240)
    assert r.stdout.count('"ok": true') == np, r.stdout[-3000:]


def test_initializer_tracing_and_timing_helpers(tmp_path):
    trace_file = tmp_path / "trace.json"
    r = launch(2, "initializer", env={"KUNGFU_ENABLE_TRACE": "1", "KUNGFU_TRACE_FILE": str(trace_file)})
    assert r.stdout.count('"ok": true') == 2 and '"traced": true' in r.stdout, r.stdout[-3000:]
    assert "[kungfu trace]" in r.stdout and "init.broadcast" in r.stdout  # per-scope summary printed at exit
    files = sorted(tmp_path.glob("trace.*.json"))  # one per worker
    assert len(files) == 2, list(tmp_path.iterdir())
    ev = json.loads(files[0].read_text())["traceEvents"]
    assert any(e["name"] == "init.broadcast" and e["ph"] == "X" and e["dur"] > 0 for e in ev)
    r = launch(3, "initializer")
    assert r.stdout.count('"ok": true') == 3 and '"traced": false' in r.stdout, r.stdout[-3000:]


def test_size_mismatch_fails_fast_with_a_message():
    t0 = time.time()
    r = launch(2, "mismatch", timeout=60, expect_rc=None)
    assert r.returncode != 0 and "chunk size mismatch" in r.stdout and "unexpected success" not in r.stdout, r.stdout[-3000:]
    assert "same shape and dtype" in r.stdout and time.time() - t0 < 40


def test_metrics_endpoint():
    r = launch(3, "metrics", env={"KUNGFU_CONFIG_ENABLE_MONITORING": "true"})
    assert r.stdout.count('"ok": true') == 3, r.stdout[-3000:]


@pytest.mark.parametrize("np", [2, 4])
def test_set_tree(np):
    r = launch(np, "set_tree")
    assert len([j for j in json_lines(r.stdout) if j.get("ok")]) == np, r.stdout[-3000:]
    assert "set_tree latency" in r.stdout


@pytest.mark.parametrize("np", [1, 2, 3, 4])
def test_optimizer_closed_forms(np):
    r = launch(np, "optimizers")
    assert len([j for j in json_lines(r.stdout) if j.get("ok")]) == np, r.stdout[-3000:]


def test_slp_result_is_np_invariant():
    """BASELINE config #1 ("SLP S-SGD world_size=2 on CPU") + the reference's golden-accuracy invariance
    (tests/python/integration/test_mnist_slp.py:149-176): same global batch => same model for every np."""
    res = {}
    for np in (1, 2, 3, 4):
        r = launch(np, "slp", extra=("--batch", "60"))
        js = [j for j in json_lines(r.stdout) if j.get("slp")]
        assert len(js) == 1, r.stdout[-2000:]
        res[np] = js[0]
    assert res[1]["acc"] > 0.5
    for np in (2, 3, 4):
        assert abs(res[np]["acc"] - res[1]["acc"]) < 1e-6, res
        assert abs(res[np]["digest"] - res[1]["digest"]) < 1e-3 * max(1.0, res[1]["digest"]), res


def test_fail_fast():
    """A failing worker cancels its siblings and the launcher exits non-zero (≙ kungfu-bad-worker)."""
    t0 = time.time()
    r = launch(3, "bad", expect_rc=None, timeout=60)
    assert r.returncode != 0 and time.time() - t0 < 30


def test_launcher_timeout():
    t0 = time.time()
    r = launch(2, "bad", flags=("-timeout", "3s"), expect_rc=None, timeout=60)
    assert r.returncode != 0 and time.time() - t0 < 30


def test_logdir_tee(tmp_path):
    r = launch(2, "p2p", flags=("-logdir", str(tmp_path)))
    files = sorted(os.listdir(tmp_path))
    assert any(f.endswith(".stdout.log") for f in files) and any(f.endswith(".stderr.log") for f in files), files
    assert any('"ok": true' in open(os.path.join(tmp_path, f)).read() for f in files if f.endswith(".stdout.log"))


def test_launch_multiprocess():
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from kungfu_b200.cmd import launch_multiprocess\n"
            "def f(rank):\n"
            "    import torch, kungfu_b200.torch as kf\n"
            "    y = kf.ops.all_reduce_fn(torch.ones(4) * (rank + 1))\n"
            "    assert torch.equal(y, torch.full((4,), 6.0)), y\n"
            "    print('mp-ok', rank, flush=True)\n"
            "if __name__ == '__main__':\n"
            "    launch_multiprocess(f, 3)\n") % ROOT
    path = os.path.join("/tmp", f"kf_mp_{os.getpid()}.py")
    open(path, "w").write(code)
    r = subprocess.run([sys.executable, path], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120,
                       env=dict(os.environ, PYTHONPATH=ROOT))
    assert r.returncode == 0 and r.stdout.count("mp-ok") == 3, r.stdout[-3000:]


def test_torchrun_env_bootstrap():
    """RANK/WORLD_SIZE/MASTER_* (no kungfu-run): the native control plane derives its peers from the env."""
    base = 29000 + (os.getpid() % 500)
    procs = []
    for rank in range(2):
        e = dict(os.environ, PYTHONPATH=ROOT, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                 MASTER_PORT=str(base))
        for k in list(e):
            if k.startswith("KUNGFU_"):
                del e[k]
        procs.append(subprocess.Popen([sys.executable, WORKER, "p2p"], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all('"ok": true' in o for o in outs)
