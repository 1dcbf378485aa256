#!/usr/bin/env python
"""Turn an .ncu-rep into the markdown summary that gets committed under profiles/ (the reports themselves are scratch).

    python benchmarks/ncu_summary.py gpurun_out/ncu_gemm_tc5.ncu-rep [--title "..."] [--top 8] >> profiles/gemm_tc5_ncu.md
Per kernel: duration, DRAM bytes and % of peak, L2 / SM / tensor-pipe utilisation, achieved occupancy, registers, launch shape, and
the SASS instructions that collect the most warp-stall samples with their dominant stall reason (needs -lineinfo / --import-source).
"""
import argparse
import csv
import io
import subprocess


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--title", default="")
    ap.add_argument("--top", type=int, default=8)
    a = ap.parse_args()
    raw = list(csv.reader(io.StringIO(ncu(["-i", a.rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    ix = {h: i for i, h in enumerate(hdr)}
    metrics = [("duration", "gpu__time_duration.sum"), ("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"),
               ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("L2 % of peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
               ("SM % of peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed"), ("tensor pipe % (active cycles)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
               ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("registers/thread", "launch__registers_per_thread"),
               ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
               ("instructions", "smsp__inst_executed.sum")]
    print(f"## {a.title or a.rep}\n")
    seen = {}
    for k, r in enumerate(raw[2:]):
        name = r[ix["Kernel Name"]]
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 1:
            continue  # first launch of each kernel
        print(f"### `{name[:110]}`\n")
        print("| metric | value |\n|---|---|")
        for label, key in metrics:
            if key in ix and r[ix[key]] not in ("", "n/a"):
                print(f"| {label} | {r[ix[key]]} {units[ix[key]]} |")
        src = list(csv.reader(io.StringIO(ncu(["-i", a.rep, "--page", "source", "--csv", "--kernel-id", f":::{k + 1}"]))))
        if len(src) > 3:
            sh = src[1]
            sx = {h: i for i, h in enumerate(sh)}
            col = "Warp Stall Sampling (All Samples)"
            if col in sx:
                rows = []
                for s in src[2:]:
                    try:
                        rows.append((float(s[sx[col]]), s))
                    except (ValueError, IndexError):
                        pass
                tot = sum(v for v, _ in rows) or 1.0
                rows.sort(key=lambda x: -x[0])
                print(f"\nTop stall sites ({int(tot)} samples):\n\n| % of samples | SASS | dominant stall |\n|---:|---|---|")
                uniq, last = [], None
                for v, s in rows:  # inlined code shows up as repeated (address, SASS) rows: keep one
                    if (s[0], s[1]) != last:
                        uniq.append((v, s))
                    last = (s[0], s[1])
                for v, s in uniq[:a.top]:
                    reasons = sorted(((float(s[sx[c]]), c) for c in sh if c.startswith("stall_") and "Not Issued" not in c and s[sx[c]] not in ("", "0")), reverse=True)
                    print(f"| {100 * v / tot:.1f} | `{s[1].strip()[:70]}` | {reasons[0][1] if reasons else '-'} |")
        print()


if __name__ == "__main__":
    main()
