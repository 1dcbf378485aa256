#!/usr/bin/env python
"""tcgen05 GEMM + bias + GELU (our CUTLASS-built kernel, epilogue out of TMEM) against cuBLAS `F.linear` followed by a separate GELU.

Shapes are the feed-forward up-projections of BERT-base / BERT-large at the benchmark batch (tokens = batch x seq) plus an
8192^3-class square. Device-timed with CUDA events after warm-up; the input set per iteration (X, W, Y) exceeds nothing
here, so L2 is flushed between iterations by writing a 256 MB buffer. Roofline: the MEASURED cuBLAS bf16 rate in
MEASURED_PEAKS.json (fallback 1500 TFLOP/s), per the profiling recipe.

    python benchmarks/gemm_bias_gelu.py [--out profiles/gemm_bias_gelu.json]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from kungfu_b200.ops.linear_act import linear_act, supported  # noqa: E402


def timed(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        flush.fill_(1)  # > L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    peak = 1500.0
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f).get("bf16_tflops", peak))
    except (OSError, ValueError):
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    from kungfu_b200.utils.timing import ClockSampler

    sampler = ClockSampler(torch.cuda.current_device()).start()  # nvidia-smi clocks / throttle reasons during the timed region
    rows = []
    for name, tokens, in_f, out_f in [("bert-base ffn-up b32xs128", 4096, 768, 3072), ("bert-large ffn-up b32xs128", 4096, 1024, 4096),
                                      ("bert-large ffn-up b128xs128", 16384, 1024, 4096), ("square 8192", 8192, 8192, 8192)]:
        x = torch.randn(tokens, in_f, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(out_f, in_f, device="cuda", dtype=torch.bfloat16) * 0.02
        b = torch.randn(out_f, device="cuda", dtype=torch.bfloat16)
        assert supported(x, w, b), "tcgen05 GEMM extension not available"
        with torch.no_grad():
            ours = linear_act(x, w, b, "gelu")
            ref = F.gelu(F.linear(x.float(), w.float(), b.float()))
            err = (ours.float() - ref).abs().max().item()
            t_ours = timed(lambda: linear_act(x, w, b, "gelu"), a.iters, flush)
            t_lib = timed(lambda: F.gelu(F.linear(x, w, b)), a.iters, flush)
            t_gemm = timed(lambda: F.linear(x, w, b), a.iters, flush)
        flops = 2.0 * tokens * in_f * out_f
        row = {"shape": name, "tokens": tokens, "in": in_f, "out": out_f, "max_abs_err_vs_fp32": round(err, 4),
               "ours_us": round(t_ours * 1e3, 1), "cublas_plus_gelu_us": round(t_lib * 1e3, 1), "cublas_gemm_only_us": round(t_gemm * 1e3, 1),
               "ours_tflops": round(flops / t_ours / 1e9, 1), "cublas_plus_gelu_tflops": round(flops / t_lib / 1e9, 1),
               "speedup": round(t_lib / t_ours, 2), "frac_of_measured_peak": round(flops / t_ours / 1e9 / peak, 3)}
        rows.append(row)
        print(json.dumps(row), flush=True)
    clocks = sampler.stop()
    print(json.dumps({"clocks": clocks}), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"peak_tflops": peak, "clocks": clocks, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
