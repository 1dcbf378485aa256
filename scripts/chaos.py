#!/usr/bin/env python
"""Random failure schedules against `kungfu-run -w -shrink-on-failure [-refill] [-auto-recover]` (host plane, CPU workers): every run picks a
cluster size, which workers die (crash or hang) and when, whether lost workers are replaced; it passes when the launcher exits 0 and all
workers of the final cluster report the same step count and bit-identical models.

    python scripts/chaos.py --runs 20 [--seed 1]
"""
import argparse
import json
import os
import random
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=20)
    ap.add_argument("--seed", type=int, default=int(time.time()))
    ap.add_argument("--keep-logs", default="")
    a = ap.parse_args()
    rng = random.Random(a.seed)
    sys.path.insert(0, ROOT)
    from kungfu_b200.cmd import kungfu_run_path

    fails = 0
    for i in range(a.runs):
        np_ = rng.randint(2, 5)
        steps = rng.randint(12, 40)
        refill = rng.random() < 0.4
        hang = (not refill) and rng.random() < 0.25
        n_crash = 1 if hang else rng.randint(1, max(1, np_ - 1))
        crashes = []
        size = np_
        t = rng.randint(0, 4)
        for _ in range(n_crash):
            if size <= 1:
                break
            crashes.append((rng.randrange(size), t))
            if not refill:
                size -= 1
            t += rng.choice([0, 0, 1, 3, 6]) if not refill else rng.randint(8, 12)
        crashes = [(r, s) for r, s in crashes if s < steps - 2]
        if not crashes:
            crashes = [(np_ - 1, 1)]
        spec = ",".join(f"{r}@{s}" for r, s in crashes)
        port = 20000 + (os.getpid() * 7 + i * 97) % 9000
        cmd = [kungfu_run_path(), "-q", "-w", "-shrink-on-failure", "-np", str(np_), "-H", f"127.0.0.1:{np_ + 2}", "-port", str(port), "-port-range",
               f"{port + 1}-{port + 60}", "-builtin-config-port", str(9600 + (os.getpid() + i) % 300)]
        env = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN")
        if refill:
            cmd += ["-refill", "-standby", "1"]
            env["KUNGFU_TEST_REFILL"] = "1"
            steps = max(steps, 40)
        if hang:
            cmd += ["-auto-recover", "3s"]
            env["KUNGFU_TEST_HANG"] = "1"
            env["KUNGFU_MONITOR_PORT"] = str(17000 + (os.getpid() + i) % 2000)
        cmd += [sys.executable, os.path.join(ROOT, "tests", "elastic_opt_worker.py"), "shrink", spec, str(steps)]
        t0 = time.time()
        try:
            r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
            out, rc = r.stdout, r.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), 124
        done = []
        for line in out.splitlines():
            if '"done": "shrink"' in line:
                done.append(json.loads(line[line.index("{"):]))
        ok = rc == 0 and done and all(d["identical"] and d["finite"] and d["step"] == steps for d in done) and len({d["final_size"] for d in done}) == 1 \
            and len(done) == done[0]["final_size"]
        print(f"run {i}: np={np_} steps={steps} crashes={spec} refill={refill} hang={hang} -> rc={rc} workers_done={len(done)} "
              f"final={done[0]['final_size'] if done else '?'} {time.time() - t0:.1f}s {'ok' if ok else 'FAIL'}", flush=True)
        if not ok:
            fails += 1
            if a.keep_logs:
                os.makedirs(a.keep_logs, exist_ok=True)
                with open(os.path.join(a.keep_logs, f"chaos_{a.seed}_{i}.log"), "w") as f:
                    f.write(" ".join(cmd) + "\n" + out)
    print(f"{a.runs - fails}/{a.runs} ok (seed {a.seed})")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
