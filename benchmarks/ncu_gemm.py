"""Tiny workload for an Nsight Compute capture of the hand-written tcgen05 GEMM (bias + GELU epilogue) at the BERT-large
feed-forward shape: ncu --set full --import-source on -k regex:gemm_tc5 -c 2 -o gpurun_out/ncu_gemm python benchmarks/ncu_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kungfu_b200.ops.linear_act import linear_act  # noqa: E402

tokens, in_f, out_f = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (4096, 1024, 4096)))
x = torch.randn(tokens, in_f, device="cuda", dtype=torch.bfloat16)
w = torch.randn(out_f, in_f, device="cuda", dtype=torch.bfloat16) * 0.02
b = torch.randn(out_f, device="cuda", dtype=torch.bfloat16)
with torch.no_grad():
    for _ in range(3):
        y = linear_act(x, w, b, "gelu")
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
