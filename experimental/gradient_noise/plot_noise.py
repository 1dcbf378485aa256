#!/usr/bin/env python
"""Extract loss / gradient-noise-scale traces from training logs and plot them.

Reference: experimental/gradient_noise/plot_noise.py (regex over worker log lines -> matplotlib). Inputs here are the logs our
tools write: JSON lines with `noise_scale` / `loss` / `img_per_s` keys (benchmarks/monitoring/benchmark.py, benchmarks/
adaptation/elastic_resize.py, with or without the launcher's `[ip.port::stdout]` prefix) and `RESULT:` lines. Output: a CSV,
an ASCII sparkline per series on stdout, and a PNG when matplotlib is importable.

    python experimental/gradient_noise/plot_noise.py run.log [--worker 127.0.0.1.10000] [--ema 0.9] [--out noise]
"""
import argparse
import csv
import json
import re
import sys

_PREFIX = re.compile(r"^(?:\x1b\[[0-9;]*m)*\[(?P<who>[^\]]*?)(?:::(?:stdout|stderr))?\](?:\x1b\[[0-9;]*m)*\s?")
_JSON = re.compile(r"\{.*\}")
_KEYS = ("noise_scale", "loss", "img_per_s", "variance")


def parse_log(path, worker=None):
    """-> {series name: [(index, value)]} in file order; `worker` filters on the launcher prefix."""
    series = {k: [] for k in _KEYS}
    with open(path, "r", errors="replace") as f:
        for line in f:
            who = None
            m = _PREFIX.match(line)
            if m:
                who, line = m.group("who"), line[m.end():]
            if worker is not None and who is not None and worker not in who:
                continue
            j = _JSON.search(line)
            if not j:
                continue
            try:
                rec = json.loads(j.group(0))
            except ValueError:
                continue
            for k in _KEYS:
                v = rec.get(k)
                if isinstance(v, (int, float)):
                    series[k].append((len(series[k]), float(v)))
    return {k: v for k, v in series.items() if v}


def ema(values, alpha):
    out, run = [], None
    for v in values:
        run = v if run is None else alpha * run + (1 - alpha) * v
        out.append(run)
    return out


def sparkline(values, width=60):
    if not values:
        return ""
    bars = " .:-=+*#%@"
    step = max(1, len(values) // width)
    pts = [sum(values[i:i + step]) / len(values[i:i + step]) for i in range(0, len(values), step)]
    lo, hi = min(pts), max(pts)
    span = (hi - lo) or 1.0
    return "".join(bars[int((p - lo) / span * (len(bars) - 1))] for p in pts)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("log")
    ap.add_argument("--worker", default=None, help="only lines from this worker (substring of the launcher prefix)")
    ap.add_argument("--ema", type=float, default=0.0, help="smooth with an exponential moving average of this decay")
    ap.add_argument("--out", default="noise", help="basename of the .csv / .png written")
    a = ap.parse_args(argv)
    series = parse_log(a.log, a.worker)
    if not series:
        print("no noise_scale / loss records found in", a.log)
        return 1
    n = max(len(v) for v in series.values())
    with open(a.out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["index"] + list(series))
        for i in range(n):
            w.writerow([i] + [series[k][i][1] if i < len(series[k]) else "" for k in series])
    for k, pts in series.items():
        vals = [v for _, v in pts]
        if a.ema > 0:
            vals = ema(vals, a.ema)
        print(f"{k:12s} n={len(vals):4d} min={min(vals):.4g} max={max(vals):.4g} last={vals[-1]:.4g}  |{sparkline(vals)}|")
    try:
        import matplotlib

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return 0
    fig, axes = plt.subplots(len(series), 1, figsize=(8, 2.6 * len(series)), squeeze=False)
    for ax, (k, pts) in zip(axes[:, 0], series.items()):
        vals = [v for _, v in pts]
        ax.plot(vals, alpha=0.35 if a.ema > 0 else 1.0, label=k)
        if a.ema > 0:
            ax.plot(ema(vals, a.ema), label=f"{k} (EMA {a.ema})")
        ax.set_ylabel(k)
        ax.legend(loc="best")
    axes[-1, 0].set_xlabel("record")
    fig.tight_layout()
    fig.savefig(a.out + ".png", dpi=120)
    print("wrote", a.out + ".csv", a.out + ".png")
    return 0


if __name__ == "__main__":
    sys.exit(main())
