"""kungfu-test-util: fault injection and manual proposals from outside a job.

Reference: tests/go/cmd/kungfu-test-util/kungfu-test-util.go:28-35 (`-kill <peer>` sends the control message `exit`, the
worker calls exit(0); `-propose <cluster.json>` sends a Stage to the runners) and kungfu-notify-start (pushes Stage v0).

    python -m kungfu_b200.cmd.test_util --kill 127.0.0.1:10001
    python -m kungfu_b200.cmd.test_util --propose cluster.json --version 3
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
    ap.add_argument("--propose", default="", help="cluster JSON file to push to its runners as a new Stage")
    ap.add_argument("--version", type=int, default=0)
    a = ap.parse_args(argv)
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
