#!/usr/bin/env python3
"""A virtual cluster on one machine, managed from outside: a standalone config server, one launcher per "node" (loopback aliases
127.0.0.1, 127.0.0.2, … — each alias is a host with its own port space), a node that joins late, and a client that keeps resizing the job.

Reference: tests/go/cmd/kungfu-cluster-manager-example/{kungfu-cluster-manager-example,cluster,ippool}.go (the same cast as docker
containers on a private subnet: config server, `kungfu-run -w -k` per node, a late-joining node started with `-init-version -1`, random
resizes) and benchmarks/adaptation/gen-compose.py:60-121. With docker the same layout is produced by
`benchmarks/adaptation/gen_compose.py`; here the nodes are plain processes so that it runs anywhere, CI included.

    python experiments/cluster_manager.py --nodes 3 --slots 2 --duration 8 --late-node-after 2
    python experiments/cluster_manager.py --nodes 2 --slots 4 -- python examples/mnist_elastic_policy/main.py --schedule ""

The default job is the native `kungfu-fake-adaptive-trainer`, which adopts whatever cluster the config server holds after every step.
"""
import argparse
import json
import os
import random
import socket
import struct
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=3, help="hosts of the virtual cluster, the last one joins late")
    ap.add_argument("--slots", type=int, default=2, help="worker slots per node")
    ap.add_argument("--init-np", type=int, default=2, help="workers the job starts with")
    ap.add_argument("--duration", type=float, default=8.0, help="seconds of random resizing")
    ap.add_argument("--period", type=float, default=1.0)
    ap.add_argument("--late-node-after", type=float, default=2.0, help="seconds after which the last node is started (0: with the others)")
    ap.add_argument("--config-port", type=int, default=0)
    ap.add_argument("--port", type=int, default=0, help="runner port on every node")
    ap.add_argument("--max-step", type=int, default=100000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--logdir", default="")
    ap.add_argument("job", nargs=argparse.REMAINDER, help="-- prog args…  (default: the native fake adaptive trainer)")
    a = ap.parse_args(argv)

    from kungfu_b200 import _build, _kfrt
    from kungfu_b200.cmd import config_client

    _build.build_runtime()
    kungfu_run, cfg_bin = str(_build.BIN / "kungfu-run"), str(_build.BIN / "kungfu-config-server")
    job = [x for x in a.job if x != "--"] or [str(_build.BIN / "kungfu-fake-adaptive-trainer"), "-max-step", str(a.max_step), "-slow-ms", "30", "-fast-ms", "10"]
    hosts = [f"127.0.0.{i + 1}" for i in range(a.nodes)]
    hostlist = ",".join(f"{h}:{a.slots}" for h in hosts)
    cfg_port = a.config_port or 9100 + os.getpid() % 400
    rport = a.port or 12000 + (os.getpid() * 13) % 15000
    prange = f"{rport + 1}-{rport + 60}"
    url = f"http://127.0.0.1:{cfg_port}/config"
    capacity = a.nodes * a.slots
    rng = random.Random(a.seed)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), KUNGFU_CONFIG_LOG_LEVEL=os.environ.get("KUNGFU_CONFIG_LOG_LEVEL", "WARN"))
    events, procs, outputs = [], {}, {}

    def note(kind, **kw):
        ev = dict(t=round(time.time() - t_start, 2), event=kind, **kw)
        events.append(ev)
        print(json.dumps(ev), flush=True)

    def start_node(i, late):
        cmd = [kungfu_run, "-q", "-w", "-k", "-np", str(a.init_np), "-H", hostlist, "-self", hosts[i], "-port", str(rport), "-port-range", prange,
               "-config-server", url]
        if late:
            cmd += ["-init-version", "-1"]  # knows nothing about the job yet: waits until the members push it a stage
        if a.logdir:
            cmd += ["-logdir", os.path.join(a.logdir, hosts[i])]
        p = subprocess.Popen(cmd + job, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        procs[i] = p
        outputs[i] = []
        threading.Thread(target=lambda: outputs[i].extend(p.stdout), daemon=True).start()
        note("node started", node=hosts[i], late=late)

    t_start = time.time()
    # 1. config server, seeded with the initial cluster (the first init_np slots in host order)
    seed_path = f"/tmp/kungfu-cluster-manager-{os.getpid()}.json"
    runners = ",".join(f"{h}:{rport}" for h in hosts)
    workers, left = [], a.init_np
    for h in hosts:
        for s in range(min(a.slots, left)):
            workers.append(f"{h}:{rport + 1 + s}")
        left = a.init_np - len(workers)
    with open(seed_path, "w") as f:
        f.write(_kfrt.cluster_json(runners, ",".join(workers)))
    cfg = subprocess.Popen([cfg_bin, "-port", str(cfg_port), "-init", seed_path], stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    rc = 1
    try:
        for _ in range(100):
            if _kfrt.http_request("GET", url, "")[0] == 200:
                break
            time.sleep(0.05)
        note("config server up", url=url, init_workers=len(workers), capacity=capacity)
        # 2. the nodes; the last one later
        late = a.nodes - 1 if a.nodes > 1 and a.late_node_after > 0 else -1
        for i in range(a.nodes):
            if i != late:
                start_node(i, late=False)
        # 3. resizes from outside; the late node appears in the middle of them
        t_end = time.time() + a.duration
        late_started = late < 0
        sizes = []
        while time.time() < t_end:
            if not late_started and time.time() - t_start >= a.late_node_after:
                start_node(late, late=True)
                late_started = True
            limit = capacity if late_started else capacity - a.slots  # do not ask for slots of a node that is not there yet
            n = rng.randint(1, max(1, limit))
            if config_client.put_size(url, n):
                sizes.append(n)
                note("resize proposed", size=n)
            time.sleep(a.period)
        if not late_started:
            start_node(late, late=True)
        # 4. wind down: shrink to one worker, let it finish, then release the idle launchers (-k keeps them alive by design)
        config_client.put_size(url, 1)
        note("resize proposed", size=1, final=True)
        time.sleep(max(1.0, 2 * a.period))
        st, body = _kfrt.http_request("GET", url, "")
        body = body.decode() if isinstance(body, (bytes, bytearray)) else body
        for w in json.loads(body).get("Workers", []) if st == 200 else []:
            # the last worker would train until its own max-step: stop it the way `kungfu-test-util -kill` does (it exits 0)
            addr = f"{socket.inet_ntoa(struct.pack('!I', int(w['IPv4'])))}:{w['Port']}"
            try:
                _kfrt.send_control(addr, "exit", "")
                note("worker stopped", worker=addr)
            except Exception as e:  # already gone
                note("worker stop failed", error=str(e))
        time.sleep(0.5)
        _kfrt.http_request("GET", f"http://127.0.0.1:{cfg_port}/stop", "")
        for i, h in enumerate(hosts):
            if i in procs and procs[i].poll() is None:
                try:
                    _kfrt.send_control(f"{h}:{rport}", "exit", "")
                except Exception:
                    pass
        for i, p in procs.items():
            try:
                p.wait(timeout=30)
            except subprocess.TimeoutExpired:
                p.kill()
        text = {hosts[i]: "".join(outputs[i]) for i in procs}
        joined = sum(t.count("sync step: 0 ->") for t in text.values())
        detached = sum(t.count("no longer in the cluster") for t in text.values())
        wrong = sum(t.count("cluster size") for t in text.values())
        late_used = late >= 0 and "sync step" in text.get(hosts[late], "")
        summary = {"summary": True, "nodes": a.nodes, "capacity": capacity, "sizes_proposed": sizes, "workers_joined": joined, "workers_detached": detached,
                   "late_node_ran_workers": bool(late_used), "wrong_steps": wrong, "node_exit_codes": {hosts[i]: p.returncode for i, p in procs.items()}}
        print(json.dumps(summary), flush=True)
        rc = 0 if wrong == 0 and all(p.returncode == 0 for p in procs.values()) else 1
        if rc:
            for h, t in text.items():
                sys.stderr.write(f"---- {h}\n{t[-3000:]}\n")
    finally:
        for p in list(procs.values()) + [cfg]:
            if p.poll() is None:
                p.kill()
        try:
            os.unlink(seed_path)
        except OSError:
            pass
    return rc


if __name__ == "__main__":
    sys.exit(main())
