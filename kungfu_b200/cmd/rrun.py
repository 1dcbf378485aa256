"""kungfu-rrun / kungfu-distribute: SSH fan-out of `kungfu-run` (or any command) to every host of a job.

Reference: srcs/go/cmd/kungfu-rrun, srcs/go/cmd/kungfu-distribute, srcs/go/utils/runner/remote/remote.go:22-131
(`RunStaticKungFuJob` / `RunElasticKungFuJob`) and srcs/go/utils/ssh. One `ssh user@host kungfu-run -self <ip> -np N -H
... prog args` per host, output prefixed per host, first failure cancels the rest.

    python -m kungfu_b200.cmd.rrun -np 16 -H 10.0.0.1:8,10.0.0.2:8 [-u user] [-w] [--dry-run] prog args...
    python -m kungfu_b200.cmd.rrun -np 4 -H 127.0.0.1:2,127.0.0.2:2 --local prog args...   # loopback-alias "hosts", no ssh

`--local` starts the per-host launchers as local processes instead of through ssh: with loopback aliases as hosts this runs a
multi-host job on one machine (every peer listens on its own alias), which is how the multi-host tier is tested here.
"""
from __future__ import annotations

import argparse
import shlex
import subprocess
import sys
import threading
from typing import List


def build_commands(np: int, hosts: str, prog: List[str], user: str = "", watch: bool = False, strategy: str = "", extra: List[str] = (),
                   kungfu_run: str = "kungfu-run", local: bool = False) -> List[List[str]]:
    """One command line per host: through ssh, or (local=True) the launcher itself for loopback-alias hosts on this machine."""
    cmds = []
    for h in hosts.split(","):
        ip = h.split(":")[0]
        inner = [kungfu_run, "-self", ip, "-np", str(np), "-H", hosts]
        if watch:
            inner.append("-w")
        if strategy:
            inner += ["-strategy", strategy]
        inner += list(extra) + list(prog)
        if local:
            cmds.append(inner)
            continue
        target = f"{user}@{ip}" if user else ip
        cmds.append(["ssh", "-o", "StrictHostKeyChecking=no", target, " ".join(shlex.quote(x) for x in inner)])
    return cmds


def distribute_commands(hosts: str, prog: List[str], user: str = "") -> List[List[str]]:
    """kungfu-distribute: run the same command on every host."""
    return [["ssh", "-o", "StrictHostKeyChecking=no", (f"{user}@" if user else "") + h.split(":")[0], " ".join(shlex.quote(x) for x in prog)]
            for h in hosts.split(",")]


def _tag(cmd: List[str]) -> str:
    return cmd[3] if cmd[0] == "ssh" else cmd[cmd.index("-self") + 1]


def run_all(cmds: List[List[str]]) -> int:
    procs = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for c in cmds]
    rc = [0] * len(procs)

    def pump(i: int) -> None:
        for line in procs[i].stdout:  # type: ignore[union-attr]
            sys.stdout.write(f"[{_tag(cmds[i])}] {line}")
        rc[i] = procs[i].wait()
        if rc[i] != 0:
            for p in procs:
                if p.poll() is None:
                    p.terminate()  # the exact children we started

    ts = [threading.Thread(target=pump, args=(i,)) for i in range(len(procs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return next((c for c in rc if c), 0)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="kungfu-rrun")
    ap.add_argument("-np", type=int, default=1)
    ap.add_argument("-H", dest="hosts", required=True)
    ap.add_argument("-u", dest="user", default="")
    ap.add_argument("-w", dest="watch", action="store_true")
    ap.add_argument("-strategy", default="")
    ap.add_argument("-port", default="", help="runner port, forwarded to kungfu-run")
    ap.add_argument("-port-range", dest="port_range", default="", help="worker port range, forwarded to kungfu-run")
    ap.add_argument("-q", dest="quiet", action="store_true", help="forwarded to kungfu-run")
    ap.add_argument("-standby", default="", help="warm standby workers per host, forwarded to kungfu-run")
    ap.add_argument("-auto-recover", dest="auto_recover", default="", help="heartbeat threshold, forwarded to kungfu-run")
    ap.add_argument("-timeout", default="", help="forwarded to kungfu-run")
    ap.add_argument("-shrink-on-failure", dest="shrink", action="store_true", help="forwarded to kungfu-run")
    ap.add_argument("-refill", action="store_true", help="forwarded to kungfu-run")
    ap.add_argument("-config-server", dest="config_server", default="", help="forwarded to kungfu-run")
    ap.add_argument("--distribute", action="store_true", help="kungfu-distribute: run the command itself on every host")
    ap.add_argument("--local", action="store_true", help="start the per-host launchers locally (loopback-alias hosts), no ssh")
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("prog", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    kr = "kungfu-run"
    if a.local:
        from . import kungfu_run_path

        kr = kungfu_run_path()
    extra = (["-port", a.port] if a.port else []) + (["-port-range", a.port_range] if a.port_range else []) + (["-q"] if a.quiet else [])
    extra += (["-standby", a.standby] if a.standby else []) + (["-auto-recover", a.auto_recover] if a.auto_recover else [])
    extra += (["-timeout", a.timeout] if a.timeout else []) + (["-shrink-on-failure"] if a.shrink else []) + (["-refill"] if a.refill else [])
    extra += ["-config-server", a.config_server] if a.config_server else []
    cmds = distribute_commands(a.hosts, a.prog, a.user) if a.distribute else build_commands(a.np, a.hosts, a.prog, a.user, a.watch, a.strategy,
                                                                                           extra=extra, kungfu_run=kr, local=a.local)
    if a.dry_run:
        for c in cmds:
            print(" ".join(shlex.quote(x) for x in c))
        return 0
    return run_all(cmds)


if __name__ == "__main__":
    sys.exit(main())
