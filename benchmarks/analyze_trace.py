#!/usr/bin/env python
"""Read the chrome traces `bench.py --trace` writes (one per rank) and attribute a training step:

  * step span per rank (first kernel of a replay -> last kernel), busy time of the compute stream, idle gaps in it
  * every collective / fused kernel of ours (or ncclDevKernel) with its start offset inside the step, its duration, and how
    long the compute stream was simultaneously busy (overlap) — i.e. which part of it is exposed
  * the tail: time between the end of backward's last compute kernel and the end of the step

    python benchmarks/analyze_trace.py gpurun_out/trace_ours2_gated [--md profiles/scale_trace.md]
"""
import argparse
import glob
import json
import re
import statistics
import sys

COMM_PAT = re.compile(r"ssgd_fused|sma_fused|sma_blend|allreduce_|barrier_kernel|wait_done|ncclDevKernel|pair_avg|pull_tma|pull_kernel|publish_kernel")


def load(path):
    d = json.load(open(path))
    evs = [e for e in d["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    evs.sort(key=lambda e: e["ts"])
    return evs


def split_steps(evs, nsteps=3):
    """The trace holds `nsteps` graph replays back to back: cut at the largest gaps between consecutive kernel starts of the
    dominant (compute) stream is fragile; instead cut by kernel-name periodicity: the first kernel name repeats every step."""
    first = evs[0]["name"]
    idx = [i for i, e in enumerate(evs) if e["name"] == first]
    # the first kernel of a step may also occur inside it; choose nsteps cut points that are evenly spaced in index
    per = len(evs) // nsteps
    cuts = [0]
    for k in range(1, nsteps):
        target = k * per
        cuts.append(min(idx, key=lambda i: abs(i - target)))
    cuts.append(len(evs))
    return [evs[cuts[i]:cuts[i + 1]] for i in range(nsteps)]


def merged_busy(intervals):
    out = []
    for s, e in sorted(intervals):
        if out and s <= out[-1][1]:
            out[-1][1] = max(out[-1][1], e)
        else:
            out.append([s, e])
    return out


def overlap_with(busy, s, e):
    t = 0.0
    for bs, be in busy:
        if be <= s:
            continue
        if bs >= e:
            break
        t += min(be, e) - max(bs, s)
    return t


def analyze_rank(path):
    evs = load(path)
    steps = split_steps(evs)
    rows = []
    for st in steps[1:]:  # skip the first replay (profiler warm-up)
        t0 = st[0]["ts"]
        t1 = max(e["ts"] + e["dur"] for e in st)
        # "compute" = every kernel that is not one of the collective / gate / wait kernels (graph replays spread the compute
        # kernels over several internal streams, so the stream id is not a usable discriminator)
        main_id = None
        main = [e for e in st if not COMM_PAT.search(e["name"])]
        busy = merged_busy([(e["ts"], e["ts"] + e["dur"]) for e in main])
        main_busy = sum(b - a for a, b in busy)
        main_end = busy[-1][1]
        comm = []
        for e in st:
            if COMM_PAT.search(e["name"]):
                s, d = e["ts"], e["dur"]
                comm.append({"name": e["name"][:60], "at_ms": (s - t0) / 1e3, "dur_us": d, "hidden_us": overlap_with(busy, s, s + d),
                             "grid": e["args"].get("grid"), "stream": e["args"].get("stream")})
        rows.append({"span_ms": (t1 - t0) / 1e3, "main_busy_ms": main_busy / 1e3, "main_idle_ms": (main_end - t0 - main_busy) / 1e3,
                     "tail_ms": (t1 - main_end) / 1e3, "n_kernels": len(st), "comm": comm})
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("prefix")
    ap.add_argument("--md", default="")
    a = ap.parse_args()
    files = sorted(glob.glob(a.prefix + ".rank*.chrome.json"))
    if not files:
        sys.exit("no trace files for " + a.prefix)
    out = [f"## {a.prefix}\n"]
    for f in files:
        rows = analyze_rank(f)
        r = rows[-1]
        span = statistics.mean(x["span_ms"] for x in rows)
        out.append(f"### {f.split('/')[-1]}: step span {span:.3f} ms, compute-stream busy {r['main_busy_ms']:.3f} ms, idle inside {r['main_idle_ms']:.3f} ms, "
                   f"tail after last compute kernel {r['tail_ms']:.3f} ms, {r['n_kernels']} kernels\n")
        out.append("| at (ms) | kernel | grid | dur (us) | hidden under compute (us) | exposed (us) |")
        out.append("|---:|---|---|---:|---:|---:|")
        tot_d = tot_e = 0.0
        for c in r["comm"]:
            exp = c["dur_us"] - c["hidden_us"]
            tot_d += c["dur_us"]
            tot_e += exp
            out.append(f"| {c['at_ms']:.3f} | `{c['name']}` | {c['grid']} | {c['dur_us']:.1f} | {c['hidden_us']:.1f} | {exp:.1f} |")
        out.append(f"\ncollective kernels: {tot_d / 1e3:.3f} ms total, {tot_e / 1e3:.3f} ms not overlapped with a compute kernel\n")
    text = "\n".join(out)
    print(text)
    if a.md:
        with open(a.md, "a") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
