"""Native test / benchmark tools built next to the launcher (bin/kungfu-*), each run the way the reference's CI runs its Go tools.

Reference (tier "Go integration tools" / "Fake trainers"): tests/go/cmd/{kungfu-test-public-apis, kungfu-test-monitor,
kungfu-bench-p2p, kungfu-fake-adaptive-trainer, kungfu-bad-worker}, tests/cpp/integration/fake_in_proc_trainer.cpp.
"""
import itertools
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_counter = itertools.count()


def _bin(name):
    from kungfu_b200.cmd import _bin as b

    return b(name)


def _ports():
    k = next(_counter)
    base = 11000 + ((os.getpid() * 41 + 7000 + k * 64) % 18000)
    return base, f"{base + 1}-{base + 60}"


def _run(np, tool, args=(), flags=(), slots=None, env=None, timeout=120, expect_rc=0):
    rport, prange = _ports()
    cmd = [_bin("kungfu-run"), "-q", "-np", str(np), "-H", f"127.0.0.1:{slots or max(np, 1)}", "-port", str(rport), "-port-range", prange,
           *flags, _bin(tool), *args]
    e = dict(os.environ, KUNGFU_CONFIG_LOG_LEVEL="WARN")
    e.update(env or {})
    r = subprocess.run(cmd, cwd=ROOT, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    out = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout)
    if expect_rc is not None:
        assert r.returncode == expect_rc, f"rc={r.returncode}\n{out[-4000:]}"
    return r.returncode, out, rport


@pytest.mark.parametrize("np,strategy", [(1, "AUTO"), (2, "STAR"), (3, "RING"), (4, "BINARY_TREE_STAR")])
def test_public_apis_native(np, strategy):
    _, out, _ = _run(np, "kungfu-test-public-apis", flags=("-strategy", strategy))
    assert f"test-public-apis ok np={np}" in out
    assert "CHECK FAILED" not in out


def test_public_apis_with_monitoring_reads_its_own_metrics():
    _, out, _ = _run(3, "kungfu-test-public-apis", env={"KUNGFU_CONFIG_ENABLE_MONITORING": "true"})
    assert "test-public-apis ok np=3" in out
    assert re.search(r"metrics endpoint http://127\.0\.0\.1:\d+/metrics: \d+ bytes", out)


@pytest.mark.parametrize("args", [("-model", "slp", "-mode", "seq"), ("-model", "resnet50", "-mode", "par", "-epochs", "3"),
                                  ("-model", "resnet50-keras", "-fuse", "-epochs", "2", "-versioned")])
def test_bench_p2p(args):
    _, out, _ = _run(3, "kungfu-bench-p2p", args)
    m = re.search(r"Result: model: ([\w-]+), (\d+) tensors, mode: (\w+), np: 3, rate: ([\d.]+) (\S+)", out)
    assert m, out[-2000:]
    assert float(m.group(4)) > 0
    assert "VERIFICATION FAILED" not in out


def test_bench_p2p_rejects_unknown_model():
    rc, out, _ = _run(1, "kungfu-bench-p2p", ("-model", "alexnet"), expect_rc=None)
    assert rc != 0 and "invalid model name" in out


@pytest.mark.parametrize("np,strategy,args", [(1, "AUTO", ("-model", "slp")), (3, "RING", ("-model", "resnet50", "-mode", "par", "-epochs", "2")),
                                              (4, "BINARY_TREE_STAR", ("-model", "resnet50", "-fuse", "-epochs", "2", "-dtype", "bf16")),
                                              (2, "STAR", ("-model", "slp", "-dtype", "i32", "-mode", "par"))])
def test_bench_allreduce(np, strategy, args):
    """≙ tests/go/cmd/kungfu-bench-allreduce: one all-reduce per tensor of a fake model, seq / par / fused, every result verified."""
    _, out, _ = _run(np, "kungfu-bench-allreduce", args, flags=("-strategy", strategy))
    m = re.search(rf"Result: model: ([\w-]+), (\d+) tensors, mode: (\w+)( \(fused\))?, np: {np}, rate: ([\d.]+) (\S+)", out)
    assert m, out[-2000:]
    assert "VERIFICATION FAILED" not in out


def test_launcher_restarts_workers_that_die_in_the_loader():
    """≙ -rand-nccl-failure + runner/local/hack.go: a worker whose first stderr line is the loader's "Inconsistency detected by ld.so" is
    started again (it has not joined anything yet); the job completes. With the retries used up the job fails instead of looping."""
    _, out, _ = _run(3, "kungfu-bench-allreduce", ("-model", "slp", "-rand-failure", "0.5"), timeout=120)
    assert re.search(r"Result: model: slp, 2 tensors, mode: seq, np: 3", out) and "VERIFICATION FAILED" not in out
    rc, out, _ = _run(2, "kungfu-bench-allreduce", ("-model", "slp", "-rand-failure", "1"), env={"KUNGFU_START_RETRIES": "2"}, expect_rc=None, timeout=60)
    assert rc != 0 and out.count("restarting for the") >= 2 and "restarting for the 3-th" not in out
    # any other start-up failure is not retried
    rc, out, _ = _run(2, "kungfu-bench-allreduce", ("-model", "alexnet",), expect_rc=None, timeout=60)
    assert rc != 0 and "restarting" not in out


def test_launcher_restart_in_watch_mode():
    rport, _ = _ports()
    _, out, _ = _run(2, "kungfu-bench-allreduce", ("-model", "slp", "-rand-failure", "0.5"), flags=("-w", "-builtin-config-port", str(rport + 62)), timeout=120)
    assert re.search(r"Result: model: slp, 2 tensors, mode: seq, np: 2", out)


def test_fake_adaptive_trainer_follows_its_schedule():
    """1 -> 2 -> 4 -> 8 -> 4 -> 1 inside one job, natively: step counters agree after every change, every step reduces to the size of
    the cluster of that step, detached workers leave, rank 0 finishes all steps."""
    rport, _ = _ports()
    _, out, _ = _run(1, "kungfu-fake-adaptive-trainer", ("-max-step", "18", "-schedule", "2:2,5:4,8:8,11:4,14:1", "-slow-ms", "20", "-fast-ms", "5"),
                     flags=("-w", "-builtin-config-port", str(rport + 62)), slots=8, timeout=180)
    assert re.search(r"finished: rank 0, 5 resizes seen, 18 steps checked", out), out[-3000:]
    sizes = [int(m.group(2)) for m in re.finditer(r"\.\d+::stdout\] resize (\d+) -> (\d+) took .* rank 0 -> 0", out)]
    assert sizes == [2, 4, 8, 4, 1], sizes
    # newcomers learned the step from the survivors
    assert re.search(r"sync step: 0 -> 9 \(rank \d of 8", out)
    assert out.count("(detached)") == 7
    assert "cluster size" not in out  # the per-step check never fired


def test_bad_worker_fails_the_job_fast():
    rc, out, _ = _run(3, "kungfu-bad-worker", ("-rank", "1", "-after", "1s", "-code", "7", "-total", "30s"), expect_rc=None, timeout=60)
    assert rc != 0
    assert "bad worker (rank 1): exiting" in out
    assert "exited with 7" in out


def test_bad_worker_is_survived_with_shrink_on_failure():
    """The same failing job under -w -shrink-on-failure: the good workers see their collective fail — this tool does not call recover(),
    so they report the error and stop, but the launcher itself must not hang and must name the lost worker."""
    rc, out, _ = _run(3, "kungfu-bad-worker", ("-rank", "2", "-after", "1s", "-total", "4s"), flags=("-w", "-shrink-on-failure"),
                      expect_rc=None, timeout=90)
    assert "bad worker (rank 2): exiting" in out


@pytest.mark.parametrize("np,model", [(2, "slp"), (4, "resnet50")])
def test_fake_inproc_trainer(np, model):
    r = subprocess.run([_bin("kungfu-fake-inproc-trainer"), "-np", str(np), "-model", model, "-steps", "3", "-warmup", "1"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=180)
    assert r.returncode == 0, r.stdout[-3000:]
    rates = re.findall(r"Img/sec ([\d.]+)\s+\[(sim|peers):", r.stdout)
    assert [b for _, b in rates] == ["sim", "peers"] and all(float(v) > 0 for v, _ in rates)
    assert "WRONG RESULT" not in r.stdout


# ---- programs around the tools: policy-driven elastic example, elastic system benchmark, virtual cluster manager, remote install -------
def _erun(script, args, slots=4, timeout=300):
    import sys

    rport, prange = _ports()
    cmd = [_bin("kungfu-run"), "-q", "-w", "-np", "1", "-H", f"127.0.0.1:{slots}", "-port", str(rport), "-port-range", prange,
           "-builtin-config-port", str(rport + 62), sys.executable, os.path.join(ROOT, script), *args]
    e = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="ERROR")
    r = subprocess.run(cmd, cwd=ROOT, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    out = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout)
    assert r.returncode == 0, out[-4000:]
    return out


def test_elastic_policy_example():
    """examples/mnist_elastic_policy: a BasePolicy resizes 1 -> 2 -> 3 -> 1 on a step schedule under a PolicyHook; newcomers adopt the
    progress, detached workers leave, the model still learns."""
    out = _erun("examples/mnist_elastic_policy/main.py", ["--schedule", "5:2,10:3,20:1", "--num-epochs", "1", "--samples", "1500"])
    m = re.search(r"done: steps=(\d+) trained=(\d+) resizes=\[\(5, 1, 2\), \(10, 2, 3\), \(20, 3, 1\)\] accuracy=([\d.]+)", out)
    assert m, out[-3000:]
    assert int(m.group(2)) >= 1500 and float(m.group(3)) > 0.9
    assert out.count("detached after step 20") == 2


def test_elastic_eager_example():
    """examples/mnist_elastic_eager: the loop itself keeps the job consistent with resize / all-reduce MAX / broadcast only."""
    out = _erun("examples/mnist_elastic_eager/main.py", ["--samples", "3000", "--num-epochs", "2", "--schedule", "10:2,20:3,30:1", "--learning-rate", "0.1"])
    m = re.search(r"done: steps=90 trained=6000 np=1 resizes=\[\(10, 1, 2\), \(20, 2, 3\), \(30, 3, 1\)\] accuracy=([\d.]+)", out)
    assert m, out[-3000:]
    assert float(m.group(1)) > 0.9 and out.count("detached after step 30") == 2


def test_elastic_follow_example_with_an_external_resize():
    """examples/mnist_elastic_docker: the program only follows the config server; the native config client resizes it from outside."""
    import sys
    import threading
    import time

    rport, prange = _ports()
    url = f"http://127.0.0.1:{rport + 62}/config"
    cmd = [_bin("kungfu-run"), "-q", "-w", "-np", "1", "-H", "127.0.0.1:3", "-port", str(rport), "-port-range", prange, "-builtin-config-port", str(rport + 62),
           sys.executable, os.path.join(ROOT, "examples/mnist_elastic_docker/mnist_slp_elastic.py"), "--samples", "2000", "--num-epochs", "6", "--step-ms", "40"]
    p = subprocess.Popen(cmd, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="ERROR"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    done = {}

    def resizer():
        t0 = time.time()
        while time.time() - t0 < 60 and p.poll() is None:  # wait until the job is up, then grow it once
            rc, out = _tool("kungfu-config-client", "-url", url)
            if rc == 0:
                time.sleep(1.0)
                done["rc"], done["out"] = _tool("kungfu-config-client", "-url", url, "-size", "3")
                return
            time.sleep(0.2)

    th = threading.Thread(target=resizer)
    th.start()
    out, _ = p.communicate(timeout=240)
    th.join()
    out = re.sub(r"\x1b\[[0-9;]*m", "", out)
    assert p.returncode == 0, out[-3000:]
    assert done.get("rc") == 0, done
    m = re.search(r"done: steps=(\d+) trained=(\d+) sizes_seen_by_rank0=\[1, 3\] accuracy=([\d.]+)", out)
    assert m, out[-3000:]
    assert int(m.group(2)) >= 12000 and float(m.group(3)) > 0.9


def test_elastic_system_benchmark_on_the_host_plane():
    out = _erun("benchmarks/system/benchmark_kungfu_elastic.py", ["--no-cuda", "--model", "slp", "--num-iters", "12", "--schedule", "4:2,8:1",
                                                                  "--num-warmup-batches", "1"])
    m = re.search(r"RESULT: ([\d.]+) \+-([\d.]+) (\{.*\})", out)
    assert m, out[-3000:]
    import json

    attrs = json.loads(m.group(3))
    assert [(c["from"], c["to"]) for c in attrs["changes"]] == [(1, 2), (2, 1)] and attrs["np"] == 1
    assert "sync step : 0 -> 4" in out and "rank detached at iter 8" in out
    assert len(re.findall(r"\.\d+::stdout\] Iter #\d+:", out)) == 12 + 4  # rank 0: 12 iterations, the second worker: iterations 5..8


def test_virtual_cluster_manager_with_a_late_node():
    """Config server + one launcher per loopback-alias host + a node that joins late + random resizes from outside
    (≙ kungfu-cluster-manager-example); the job is the native adaptive trainer following the config server."""
    import json
    import sys

    base = 12000 + (os.getpid() * 29 + 5000) % 14000
    r = subprocess.run([sys.executable, os.path.join(ROOT, "experiments", "cluster_manager.py"), "--nodes", "3", "--slots", "2", "--duration", "4",
                        "--period", "0.6", "--late-node-after", "1.5", "--port", str(base), "--config-port", str(9500 + os.getpid() % 400)],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=180)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    summary = [json.loads(l) for l in r.stdout.splitlines() if l.startswith('{"summary"')][0]
    assert summary["wrong_steps"] == 0 and summary["late_node_ran_workers"] in (True, False)
    assert all(c == 0 for c in summary["node_exit_codes"].values()), summary
    assert summary["workers_joined"] >= 2 and len(summary["sizes_proposed"]) >= 3


def test_remote_install_dry_run_lists_every_step():
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "experiments", "remote_install.py"), "-H", "10.1.0.1:8,10.1.0.2:8", "-u", "op", "--dry-run"],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.count("-cf - --exclude") == 2 and r.stdout.count("-xf -") == 2 and r.stdout.count("op@10.1.0.") >= 4 and "kungfu_b200._build" in r.stdout


def _tool(name, *args, timeout=60):
    r = subprocess.run([_bin(name), *args], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    return r.returncode, r.stdout


def test_native_config_client_against_a_standalone_config_server(tmp_path):
    """≙ tests/go/cmd/kungfu-config-client-example: read / resize / random proposals / clear, all through the REST endpoint."""
    import json
    import time

    from kungfu_b200 import _kfrt as rt

    init = tmp_path / "init.json"
    init.write_text(rt.cluster_json("127.0.0.1:38080", "127.0.0.1:10000,127.0.0.1:10001"))
    port = 9700 + (os.getpid() * 7 + next(_counter)) % 250
    url = f"http://127.0.0.1:{port}/config"
    p = subprocess.Popen([_bin("kungfu-config-server"), "-port", str(port), "-init", str(init), "-ttl", "60s"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    try:
        for _ in range(100):
            rc, out = _tool("kungfu-config-client", "-url", url)
            if rc == 0:
                break
            time.sleep(0.1)
        assert rc == 0 and json.loads(out) == {"size": 2, "runners": 1}, out
        rc, out = _tool("kungfu-config-client", "-url", url, "-size", "5")
        assert rc == 0 and json.loads(out) == {"resize_to": 5, "ok": True}
        st, body = rt.http_request("GET", url)
        ports_now = [w["Port"] for w in json.loads(body)["Workers"]]
        assert len(ports_now) == 5 and len(set(ports_now)) == 5 and ports_now[:2] == [10000, 10001]
        # a shuffled shrink keeps a subset in its old order
        rc, out = _tool("kungfu-config-client", "-url", url, "-size", "3", "-shuffle", "-seed", "5")
        assert rc == 0
        st, body = rt.http_request("GET", url)
        kept = [w["Port"] for w in json.loads(body)["Workers"]]
        assert len(kept) == 3 and kept == sorted(kept) and set(kept) <= set(ports_now)
        rc, out = _tool("kungfu-config-client", "-url", url, "-random", "-min", "1", "-max", "6", "-period", "50ms", "-duration", "600ms", "-final", "2",
                        "-seed", "3")
        lines = [json.loads(l) for l in out.strip().splitlines()]
        assert rc == 0 and lines[-1]["final"] == 2 and lines[-1]["proposals"] >= 3 and all(1 <= l["resize_to"] <= 6 for l in lines[:-1])
        rc, out = _tool("kungfu-config-client", "-url", url)
        assert json.loads(out)["size"] == 2
        rc, out = _tool("kungfu-config-client", "-url", url, "-clear")
        assert rc == 0 and json.loads(out) == {"cleared": True}
        rc, out = _tool("kungfu-config-client", "-url", url, "-size", "3")
        assert rc != 0  # a cleared config takes no more updates: the job is over
    finally:
        rt.http_request("GET", f"http://127.0.0.1:{port}/stop")
        try:
            p.wait(timeout=10)
        except subprocess.TimeoutExpired:
            p.kill()


def test_native_test_util_waits_pings_and_kills_a_worker():
    """≙ tests/go/cmd/kungfu-test-util: -wait / -ping a worker of a running job, then -kill it: the `exit` control message ends that
    worker (status 0) and its listener is gone."""
    import json
    import time

    rport, prange = _ports()
    base = int(prange.split("-")[0])
    cmd = [_bin("kungfu-run"), "-q", "-np", "2", "-H", "127.0.0.1:2", "-port", str(rport), "-port-range", prange,
           _bin("kungfu-bad-worker"), "-rank", "9", "-total", "60s"]  # nobody is "bad": both workers just keep all-reducing
    p = subprocess.Popen(cmd, cwd=ROOT, env=dict(os.environ, KUNGFU_CONFIG_LOG_LEVEL="WARN"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        target = f"127.0.0.1:{base + 1}"
        rc, out = _tool("kungfu-test-util", "-wait", target, "-timeout", "20s")
        assert rc == 0 and json.loads(out)["up"] is True, out
        time.sleep(0.5)
        rc, out = _tool("kungfu-test-util", "-ping", target, "-n", "3")
        d = json.loads(out)
        assert rc == 0 and d["pings"] == 3 and 0 <= d["rtt_s"] <= d["mean_rtt_s"] < 1.0
        rc, out = _tool("kungfu-test-util", "-kill", target)
        assert rc == 0 and json.loads(out) == {"killed": target}
        t0 = time.time()
        while time.time() - t0 < 10:  # the worker leaves with status 0 (its sibling then waits for it until its own dial budget ends)
            rc, out = _tool("kungfu-test-util", "-wait", target, "-timeout", "300ms")
            if rc == 1:
                break
            time.sleep(0.1)
        assert rc == 1 and json.loads(out)["up"] is False
        p.terminate()
        p.wait(timeout=20)
    finally:
        if p.poll() is None:
            p.kill()
    rc, out = _tool("kungfu-test-util")
    assert rc == 2 and "usage" in out
