"""kungfu-test-util: fault injection and manual proposals from outside a job.

Reference: tests/go/cmd/kungfu-test-util/kungfu-test-util.go:28-35 (`-kill <peer>` sends the control message `exit`, the
worker calls exit(0); `-propose <cluster.json>` sends a Stage to the runners) and kungfu-notify-start (pushes Stage v0).

    python -m kungfu_b200.cmd.test_util --kill 127.0.0.1:10001
    python -m kungfu_b200.cmd.test_util --propose cluster.json --version 3
    python -m kungfu_b200.cmd.test_util --wait 127.0.0.1:10001 --timeout 10

The native twin is `bin/kungfu-test-util` (csrc/launcher/kungfu_test_util.cpp), same operations with the Go tools' `-flag value` grammar.
"""
from __future__ import annotations

import argparse
import json
import sys


def main(argv=None) -> int:
    from .. import _kfrt  # type: ignore[attr-defined]

    ap = argparse.ArgumentParser(prog="kungfu-test-util")
    ap.add_argument("--kill", default="", help="ip:port of the worker (or runner) to stop")
    ap.add_argument("--ping", default="")
    ap.add_argument("--wait", default="", help="return once the listener at ip:port answers (exit 1 after --timeout seconds)")
    ap.add_argument("--timeout", type=float, default=10.0)
    ap.add_argument("--notify-start", default="", help="cluster JSON file pushed to its runners as Stage v0 (runners started with -init-version -1)")
    ap.add_argument("--propose", default="", help="cluster JSON file to push to its runners as a new Stage")
    ap.add_argument("--version", type=int, default=0)
    a = ap.parse_args(argv)
    if a.wait:
        import time

        t0, up = time.time(), False
        while not up and time.time() - t0 < a.timeout:
            up = _kfrt.ping(a.wait) >= 0
            if not up:
                time.sleep(0.05)
        print(json.dumps({"peer": a.wait, "up": up, "waited_s": round(time.time() - t0, 3)}))
        if not up:
            return 1
    if a.notify_start:
        a.propose, a.version = a.notify_start, 0
    if a.kill:
        _kfrt.send_control(a.kill, "exit", "")
    if a.ping:
        print(json.dumps({"peer": a.ping, "rtt_s": _kfrt.ping(a.ping)}))
    if a.propose:
        cluster = open(a.propose).read()
        stage = _kfrt.stage_json(a.version, cluster)
        for r in _kfrt.cluster_from_json(cluster)["runners"]:
            _kfrt.send_control(r, "update", stage)
    return 0


if __name__ == "__main__":
    sys.exit(main())
