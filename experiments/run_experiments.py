#!/usr/bin/env python3
"""Experiment sweep driver: runs a grid of (job template x world size x optimizer x model) on this box — or, with --hosts, on a
cluster through kungfu-rrun — and collects the RESULT: lines into one JSON/markdown table.

Reference: experiments/ (Go drivers that SSH-launch sweeps: `kungfu-run-scaling-experiments`, `kungfu-run-train-experiments`, job
templates `elastic/elastic.go`, `tfkeras/tf-keras.go`: MobileNetV2 / ResNet50 x sync-sgd / sync-sgd-nccl).

    python experiments/run_experiments.py scaling --np 1,2,4,8 --optimizers sync-sgd,sma,async-sgd --models resnet50 [--dry-run]
    python experiments/run_experiments.py baselines --np 8 --methods ddp,ps
    python experiments/run_experiments.py elastic --schedule 4:10,8:10,4:10
    python experiments/run_experiments.py scaling --hosts 10.0.0.1:8,10.0.0.2:8 --np 16 -u me          # cluster, via kungfu-rrun
"""
import argparse
import itertools
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def torchrun(np, script, args):
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={np}", "--master-addr", "127.0.0.1", "--master-port",
            str(29500 + (os.getpid() + np) % 400), os.path.join(ROOT, script)] + args


def rrun(hosts, np, user, script, args):
    from kungfu_b200 import _build

    return [str(_build.BIN / "kungfu-rrun"), "-np", str(np), "-H", hosts] + (["-u", user] if user else []) + [sys.executable, os.path.join(ROOT, script)] + args


def jobs(a):
    nps = [int(x) for x in a.np.split(",")]
    if a.kind == "scaling":
        for n, o, m in itertools.product(nps, a.optimizers.split(","), a.models.split(",")):
            yield {"np": n, "optimizer": o, "model": m}, ("benchmarks/system/benchmark_kungfu_torch.py", ["--kf-optimizer", o, "--model", m, "--batch-size", str(a.batch_size)])
    elif a.kind == "baselines":
        for n, me, m in itertools.product(nps, a.methods.split(","), a.models.split(",")):
            yield {"np": n, "method": me, "model": m}, ("benchmarks/system/benchmark_baselines_torch.py", ["--method", me, "--model", m, "--batch-size", str(a.batch_size)])
    elif a.kind == "comm":
        for n, m in itertools.product(nps, a.models.split(",")):
            yield {"np": n, "model": m}, ("benchmarks/fake_trainer.py", ["--model", m])
    elif a.kind == "elastic":
        yield {"schedule": a.schedule}, ("benchmarks/adaptation/elastic_resize.py", ["--schedule", a.schedule])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kind", choices=["scaling", "baselines", "comm", "elastic"])
    ap.add_argument("--np", default="1,2,4,8")
    ap.add_argument("--optimizers", default="sync-sgd,sma,async-sgd")
    ap.add_argument("--methods", default="ddp,ps")
    ap.add_argument("--models", default="resnet50")
    ap.add_argument("--batch-size", type=int, default=256)
    ap.add_argument("--schedule", default="4:10,8:10,4:10")
    ap.add_argument("--hosts", default="", help="ip:slots,... -> launch through kungfu-rrun (one kungfu-run per host)")
    ap.add_argument("-u", dest="user", default="")
    ap.add_argument("--timeout", type=int, default=1200)
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("--out", default="experiments/results.json")
    a = ap.parse_args()
    results = []
    for meta, (script, args) in jobs(a):
        n = meta.get("np", 1)
        if a.kind == "elastic":
            from kungfu_b200 import _build

            first = int(a.schedule.split(":")[0])
            cmd = [str(_build.BIN / "kungfu-run"), "-w", "-np", str(first), "-H", "127.0.0.1:8", "-builtin-config-port", "9100", sys.executable, os.path.join(ROOT, script)] + args
        elif a.hosts:
            cmd = rrun(a.hosts, n, a.user, script, args)
        else:
            cmd = torchrun(n, script, args)
        if a.dry_run:
            print(" ".join(cmd))
            continue
        t0 = time.time()
        r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=a.timeout)
        lines = [l for l in r.stdout.splitlines() if "RESULT:" in l or l.startswith("{")]
        m = re.search(r"RESULT:\s*([0-9.]+)", r.stdout)
        results.append({**meta, "rc": r.returncode, "result": float(m.group(1)) if m else None, "lines": lines[-6:], "wall_s": round(time.time() - t0, 1)})
        print(json.dumps(results[-1]), flush=True)
    if results:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(results, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
