"""kungfu-config-client: drive an elastic job from outside by rewriting the cluster config on the config server.

Reference: tests/go/cmd/kungfu-config-client-example (every second PUTs a random-size subset of the workers,
kungfu-config-client-example.go:96-116) — the chaos half of the reference's docker-compose elastic test. Also usable by hand:

    python -m kungfu_b200.cmd.config_client --url http://127.0.0.1:9100/config --size 4        # one resize
    python -m kungfu_b200.cmd.config_client --url … --random --min 1 --max 8 --period 1 --duration 30 --final 2
"""
from __future__ import annotations

import argparse
import json
import random
import sys
import time


def put_size(url: str, size: int) -> bool:
    """GET the current cluster, resize it (shrink by truncation, grow on the least loaded runner) and PUT it back."""
    from .. import _kfrt  # type: ignore[attr-defined]

    st, body = _kfrt.http_request("GET", url, "")
    if st != 200:
        return False
    body = body.decode() if isinstance(body, (bytes, bytearray)) else body
    try:
        new = _kfrt.cluster_resize(body, int(size))
    except Exception:
        return False
    st, _ = _kfrt.http_request("PUT", url, new)
    return st == 200


def current_size(url: str) -> int:
    from .. import _kfrt  # type: ignore[attr-defined]

    st, body = _kfrt.http_request("GET", url, "")
    if st != 200:
        return -1
    body = body.decode() if isinstance(body, (bytes, bytearray)) else body
    return len(json.loads(body).get("Workers", []))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="kungfu-config-client")
    ap.add_argument("--url", required=True, help="config server endpoint, e.g. http://127.0.0.1:9100/config")
    ap.add_argument("--size", type=int, default=0, help="resize once to this many workers")
    ap.add_argument("--random", action="store_true", help="keep proposing random sizes")
    ap.add_argument("--min", type=int, default=1)
    ap.add_argument("--max", type=int, default=4)
    ap.add_argument("--period", type=float, default=1.0)
    ap.add_argument("--duration", type=float, default=10.0)
    ap.add_argument("--final", type=int, default=0, help="size to leave behind when the random phase ends")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)
    if a.size > 0:
        ok = put_size(a.url, a.size)
        print(json.dumps({"resize_to": a.size, "ok": ok}), flush=True)
        return 0 if ok else 1
    if a.random:
        rng = random.Random(a.seed)
        t_end = time.time() + a.duration
        n_put = 0
        while time.time() < t_end:
            n = rng.randint(a.min, a.max)
            if put_size(a.url, n):
                n_put += 1
                print(json.dumps({"resize_to": n}), flush=True)
            time.sleep(a.period)
        if a.final > 0:
            put_size(a.url, a.final)
        print(json.dumps({"proposals": n_put, "final": a.final or None}), flush=True)
        return 0
    print(json.dumps({"size": current_size(a.url)}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
