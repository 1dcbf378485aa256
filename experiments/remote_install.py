#!/usr/bin/env python3
"""Install this framework on every host of a cluster: ship the source tree, build the native parts there, check the result.

Reference: experiments/cmd/kungfu-remote-install/kungfu-remote-install.go (SSH to each host, fetch the sources, `pip install`, in
parallel over the host list). Here the tree is shipped as a tar stream over ssh (no git / index access is assumed on the hosts), the
native build is `python -m kungfu_b200._build` run through `kungfu-distribute` (one ssh per host, prefixed logs), and the check imports
the package and prints the versions of what was built.

    python experiments/remote_install.py -H 10.0.0.1,10.0.0.2 -u me --prefix /opt/kungfu_b200 [--python python3] [--dry-run]
    python experiments/remote_install.py -H 127.0.0.1 --local --prefix /tmp/kf-install          # no ssh: a local dry run of every step
"""
import argparse
import concurrent.futures as cf
import os
import shlex
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EXCLUDES = [".git", "gpurun_out", "baseline/_ref", "kungfu_b200/build", "kungfu_b200/bin", "kungfu_b200/lib", "*.so", "__pycache__", ".pytest_cache", "*.ncu-rep"]


def ship(host, user, prefix, local, dry):
    tar = ["tar", "-C", ROOT, "-cf", "-"] + [f"--exclude={e}" for e in EXCLUDES] + ["."]
    unpack = f"mkdir -p {shlex.quote(prefix)} && tar -C {shlex.quote(prefix)} -xf -"
    dst = ["sh", "-c", unpack] if local else ["ssh", "-o", "StrictHostKeyChecking=no", "-o", "BatchMode=yes", (f"{user}@" if user else "") + host, unpack]
    if dry:
        return 0, " ".join(map(shlex.quote, tar)) + " | " + " ".join(map(shlex.quote, dst))
    p1 = subprocess.Popen(tar, stdout=subprocess.PIPE)
    p2 = subprocess.run(dst, stdin=p1.stdout, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    p1.stdout.close()
    return (p1.wait() or p2.returncode), p2.stdout


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("-H", "--hosts", required=True, help="ip[:slots],…")
    ap.add_argument("-u", "--user", default="")
    ap.add_argument("--prefix", default="/opt/kungfu_b200")
    ap.add_argument("--python", default="python3")
    ap.add_argument("--local", action="store_true", help="run the per-host steps on this machine without ssh")
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args(argv)
    from kungfu_b200 import _build

    _build.build_runtime()
    hosts = [h.split(":")[0] for h in a.hosts.split(",") if h]
    with cf.ThreadPoolExecutor(max_workers=min(16, len(hosts))) as ex:
        for host, (rc, out) in zip(hosts, ex.map(lambda h: ship(h, a.user, a.prefix, a.local, a.dry_run), hosts)):
            print(f"[{host}] shipped to {a.prefix}: rc={rc}" + (f"\n{out}" if out else ""), flush=True)
            if rc:
                return rc
    remote = (f"cd {shlex.quote(a.prefix)} && {a.python} -m kungfu_b200._build && "
              f"{a.python} -c 'import kungfu_b200, torch; print(\"kungfu_b200 ok, torch\", torch.__version__)' && ls kungfu_b200/bin")
    dist = [str(_build.BIN / "kungfu-distribute"), "-H", ",".join(hosts)] + (["-u", a.user] if a.user else []) + (["-dry-run"] if a.dry_run else [])
    if a.local:
        cmd = ["sh", "-c", remote]
        if a.dry_run:
            print(" ".join(map(shlex.quote, cmd)))
            return 0
    else:
        cmd = dist + ["sh", "-c", remote]
    return subprocess.call(cmd)


if __name__ == "__main__":
    sys.exit(main())
