"""Multi-host jobs without a cluster: every "host" is a loopback alias (127.0.0.1, 127.0.0.2, ...) with its own `kungfu-run`,
its own workers and the same per-host port numbering; peers listen on their own alias, so the hosts do not collide.

The reference tests this tier with docker-compose on a private subnet (tests/go/cmd/kungfu-cluster-manager-example,
scripts/tests gen-compose.py; SURVEY.md §4 "Multi-host without a cluster"). Here it is ordinary processes.
"""
import os
import subprocess
import sys
import threading

import pytest

from test_runtime_multiproc import ROOT, WORKER, json_lines, kungfu_run, ports


def launch_hosts(slots, what, extra=(), flags=(), timeout=240):
    """slots = workers per host, e.g. [2, 2] -> 127.0.0.1:2,127.0.0.2:2. One launcher per host, started together."""
    hosts = [f"127.0.0.{i + 1}" for i in range(len(slots))]
    hostlist = ",".join(f"{h}:{s}" for h, s in zip(hosts, slots))
    rport, prange = ports()
    env = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN")
    outs = [None] * len(hosts)

    def run(i):
        cmd = [kungfu_run(), "-q", "-np", str(sum(slots)), "-H", hostlist, "-self", hosts[i], "-port", str(rport), "-port-range", prange,
               *flags, sys.executable, WORKER, what, *extra]
        outs[i] = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)

    ths = [threading.Thread(target=run, args=(i,)) for i in range(len(hosts))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    text = "\n".join(o.stdout for o in outs)
    assert all(o.returncode == 0 for o in outs), text[-6000:]
    return text


@pytest.mark.parametrize("slots,strategy", [([2, 2], "BINARY_TREE_STAR"), ([2, 1, 1], "TREE"),
                                            pytest.param([1, 3], "MULTI_BINARY_TREE_STAR", marks=pytest.mark.extended),
                                            pytest.param([2, 2], "RING", marks=pytest.mark.extended)])
def test_collectives_across_hosts(slots, strategy):
    out = launch_hosts(slots, "collectives", flags=["-strategy", strategy])
    assert len([j for j in json_lines(out) if j.get("ok")]) == sum(slots), out[-4000:]


@pytest.mark.parametrize("slots", [[2, 2], [3, 1], [1, 1, 2]])
def test_local_and_cross_collectives(slots):
    out = launch_hosts(slots, "multihost")
    js = [j for j in json_lines(out) if j.get("ok")]
    assert len(js) == sum(slots) and all(j["hosts"] == len(slots) for j in js), out[-4000:]
    assert sorted(j["local"][1] for j in js) == sorted(s for s in slots for _ in range(s))


def test_p2p_and_set_tree_across_hosts():
    out = launch_hosts([2, 2], "p2p")
    assert len([j for j in json_lines(out) if j.get("ok")]) == 4, out[-4000:]
    out = launch_hosts([2, 2], "set_tree")
    assert out.count('"ok": true') == 4 and "set_tree latency" in out, out[-4000:]


def test_randomised_soak_across_hosts():
    out = launch_hosts([2, 2], "stress", extra=["100"], flags=["-strategy", "BINARY_TREE_STAR"])
    assert out.count('"ok": true') == 4, out[-4000:]


@pytest.mark.extended
def test_training_result_does_not_depend_on_the_host_layout():
    """S-SGD on the SLP: 4 workers on one host and 2 + 2 on two hosts reach the same parameters (np-invariance test across
    layouts)."""
    from test_runtime_multiproc import launch

    one = [j for j in json_lines(launch(4, "slp").stdout) if j.get("slp")]
    two = [j for j in json_lines(launch_hosts([2, 2], "slp")) if j.get("slp")]
    assert len(one) == 1 and len(two) == 1
    assert one[0]["acc"] == two[0]["acc"] and abs(one[0]["digest"] - two[0]["digest"]) < 1e-4, (one[0], two[0])


def test_elastic_resize_across_hosts():
    """2 workers on host A grow to 4 (two of them on host B, whose launcher was idling in `-w -k` mode), shrink to 3, then to 1.
    Host A runs the builtin config server, host B points at it; newcomers on B inherit the state, the job result is intact."""
    import time

    from kungfu_b200 import _kfrt

    rport, prange = ports()
    cfg = 9100 + os.getpid() % 300
    hostlist = "127.0.0.1:2,127.0.0.2:2"
    env = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN")
    work = [sys.executable, WORKER, "elastic", "--schedule", "2:3,4:3,3:3,1:3", "--max-step", "12"]
    base = [kungfu_run(), "-q", "-w", "-np", "2", "-H", hostlist, "-port", str(rport), "-port-range", prange]
    b = subprocess.Popen(base + ["-k", "-self", "127.0.0.2", "-config-server", f"http://127.0.0.1:{cfg}/config"] + work, cwd=ROOT, env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        a = subprocess.run(base + ["-self", "127.0.0.1", "-builtin-config-port", str(cfg)] + work, cwd=ROOT, env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=240)
        assert a.returncode == 0, a.stdout[-5000:]
        final = [j for j in json_lines(a.stdout) if j.get("elastic")]
        assert len(final) == 1 and final[0]["final_step"] == 12 and final[0]["sizes"] == [1, 2, 3, 4] and final[0]["resizes"] == 3, a.stdout[-3000:]
        time.sleep(0.5)
        _kfrt.send_control(f"127.0.0.2:{rport}", "exit", "")  # host B's launcher idles by design (-k): tell it to leave
        out_b, _ = b.communicate(timeout=30)
        assert len([j for j in json_lines(out_b) if j.get("detached")]) == 2, out_b[-3000:]  # B's two workers came and went
    finally:
        if b.poll() is None:
            b.kill()


def test_shrink_and_continue_across_hosts():
    """Two hosts with two workers each under `-w -shrink-on-failure`; a worker on the SECOND host dies inside step 4. Its runner announces
    the loss to the survivors on both hosts and to the other runner; the three survivors regroup and finish with identical models; both
    launchers exit 0."""
    rport, prange = ports()
    cfg = 9100 + (os.getpid() + 150) % 300
    env = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN")
    work = [sys.executable, os.path.join(ROOT, "tests", "elastic_opt_worker.py"), "shrink", "3@4", "12"]
    base = [kungfu_run(), "-q", "-w", "-shrink-on-failure", "-np", "4", "-H", "127.0.0.1:2,127.0.0.2:2", "-port", str(rport), "-port-range", prange]
    b = subprocess.Popen(base + ["-self", "127.0.0.2", "-config-server", f"http://127.0.0.1:{cfg}/config"] + work, cwd=ROOT, env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        a = subprocess.run(base + ["-self", "127.0.0.1", "-builtin-config-port", str(cfg)] + work, cwd=ROOT, env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=150)
        out_b, _ = b.communicate(timeout=60)
        assert a.returncode == 0 and b.returncode == 0, (a.stdout[-3000:], out_b[-3000:])
        done = [j for j in json_lines(a.stdout + out_b) if j.get("done") == "shrink"]
        assert len(done) == 3 and all(d["identical"] and d["step"] == 12 and d["final_size"] == 3 and d["failures"] == [[4, 3]] for d in done), done
        assert "lost, continuing with 3 workers" in out_b and "lost, continuing" not in a.stdout  # the runner that owned the worker announces
    finally:
        if b.poll() is None:
            b.kill()


def test_failure_recovery_across_hosts(tmp_path):
    """One worker per host; the worker on the second host hangs inside a batch. The main monitor server (first host) notices
    the stall, tells the other host's server (`otherdown:<epochs>`), both launchers kill and restart their workers with the
    remaining epochs, and the job finishes (reference: runner/monitorserver, G17 — there it is never tested)."""
    rport, prange = ports()
    mon = 17000 + os.getpid() % 2000
    env = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN", KUNGFU_MONITOR_PORT=str(mon))
    outs = [None, None]

    def run(i):
        cmd = [kungfu_run(), "-q", "-np", "2", "-H", "127.0.0.1:1,127.0.0.2:1", "-self", f"127.0.0.{i + 1}", "-port", str(rport), "-port-range", prange,
               "-auto-recover", "3s", sys.executable, WORKER, "recovery", "--dir", str(tmp_path), "--n-epochs", "3"]
        outs[i] = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)

    ths = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    text = "\n".join(o.stdout for o in outs)
    assert all(o.returncode == 0 for o in outs), text[-5000:]
    assert all("worker failure detected" in o.stdout for o in outs), text[-5000:]  # both hosts restarted
    js = [j for j in json_lines(text) if "recovered" in j]
    assert js and js[-1] == {"recovered": True, "epochs_run_now": 2, "start_epoch": 1}, text[-5000:]


def test_rrun_local_fans_out_one_launcher_per_alias():
    """`kungfu-rrun --local`: the per-host launchers the ssh fan-out would start, run as local processes on loopback aliases."""
    env = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN")
    rport, prange = ports()
    r = subprocess.run([sys.executable, "-m", "kungfu_b200.cmd.rrun", "-np", "4", "-H", "127.0.0.1:2,127.0.0.2:2", "--local", "-q", "-port", str(rport),
                        "-port-range", prange, sys.executable,
                        os.path.join(ROOT, "examples", "torch_simple_example.py")], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=180)
    assert r.returncode == 0, r.stdout[-3000:]
    for k in range(4):
        assert f"rank {k}/4: [4.0" in r.stdout, r.stdout[-3000:]
    assert "[127.0.0.2] " in r.stdout and "[127.0.0.1] " in r.stdout


@pytest.mark.parametrize("slots,strategy", [([2, 2], "BINARY_TREE_STAR"), ([2, 1, 1], "TREE")])
def test_optimizer_closed_forms_across_hosts(slots, strategy):
    """S-SGD / SMA / PairAveraging closed forms, noise-scale / variance monitors, AdaptiveSGD switch — with the workers spread over
    several hosts."""
    out = launch_hosts(slots, "optimizers", flags=["-strategy", strategy])
    assert out.count('"ok": true') == sum(slots), out[-4000:]


def test_hostfile_and_nic_discovery(tmp_path):
    """OpenMPI-style hostfile instead of -H; the first launcher finds its own address through `-nic lo`, the second through -self."""
    hf = tmp_path / "hosts.txt"
    hf.write_text("# two hosts\n127.0.0.1 slots=2\n127.0.0.2 slots=2\n")
    rport, prange = ports()
    env = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN")
    base = [kungfu_run(), "-q", "-np", "4", "-hostfile", str(hf), "-port", str(rport), "-port-range", prange]
    prog = [sys.executable, os.path.join(ROOT, "examples", "torch_simple_example.py")]
    b = subprocess.Popen(base + ["-self", "127.0.0.2"] + prog, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    a = subprocess.run(base + ["-nic", "lo"] + prog, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
    out_b, _ = b.communicate(timeout=60)
    assert a.returncode == 0 and b.returncode == 0, a.stdout[-2000:] + out_b[-2000:]
    text = a.stdout + out_b
    for k in range(4):
        assert f"rank {k}/4: [4.0" in text, text[-3000:]


def test_native_rrun_fans_out_to_every_host_and_fails_fast():
    """`kungfu-rrun` (C++: fork/exec + pipes + poll; the reference's is Go over an SSH library) starts one launcher per host —
    here with -local on loopback-alias hosts — prefixes their output, and cancels the other hosts when one fails."""
    from kungfu_b200 import _build

    _build.build_runtime()
    rrun = str(_build.BIN / "kungfu-rrun")
    rport, prange = ports()
    env = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN")
    r = subprocess.run([rrun, "-local", "-q", "-np", "4", "-H", "127.0.0.1:2,127.0.0.2:2", "-kungfu-run", kungfu_run(), "-port", str(rport),
                        "-port-range", prange, sys.executable, WORKER, "collectives"], cwd=ROOT, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-4000:]
    assert r.stdout.count('"ok": true') == 4 and "[127.0.0.1]" in r.stdout and "[127.0.0.2]" in r.stdout, r.stdout[-4000:]
    # dry run prints the SSH command lines of a real multi-host launch
    d = subprocess.run([rrun, "-dry-run", "-np", "16", "-H", "10.0.0.1:8,10.0.0.2:8", "-u", "me", "-w", "-config-server", "http://10.0.0.1:9100/config",
                        "python3", "train.py"], stdout=subprocess.PIPE, text=True)
    lines = d.stdout.strip().splitlines()
    assert len(lines) == 2 and all(l.startswith("ssh ") and "kungfu-run -self 10.0.0." in l and " -w " in l for l in lines), d.stdout
    # the self-healing flags travel to every host's launcher; the first host runs the config server, the others are pointed at it
    d = subprocess.run([rrun, "-dry-run", "-np", "16", "-H", "10.0.0.1:8,10.0.0.2:8", "-w", "-builtin-config-port", "9100", "-standby", "2",
                        "-shrink-on-failure", "-refill", "-auto-recover", "30s", "python3", "train.py"], stdout=subprocess.PIPE, text=True)
    lines = d.stdout.strip().splitlines()
    assert len(lines) == 2 and all(" -standby 2 " in l and " -shrink-on-failure " in l and " -refill " in l and " -auto-recover 30s " in l for l in lines), d.stdout
    assert " -builtin-config-port 9100 " in lines[0] and "-config-server" not in lines[0], lines[0]
    assert " -config-server http://10.0.0.1:9100/config " in lines[1] and "-builtin-config-port" not in lines[1], lines[1]
    # kungfu-distribute: the same binary under its second name; one failing host cancels the rest and the exit code is non-zero
    dist = str(_build.BIN / "kungfu-distribute")
    f = subprocess.run([dist, "-local", "-H", "127.0.0.1,127.0.0.2", "sh", "-c", "exit 3"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=60)
    assert f.returncode == 1 and "cancelling" in f.stdout
