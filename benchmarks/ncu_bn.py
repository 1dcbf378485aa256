#!/usr/bin/env python
"""Driver for an `ncu --set full` capture of the fused BN / max-pool kernels at their largest ResNet-50 shapes (batch 256):
BN(+residual)+ReLU over [256, 256, 56, 56] (205 M elements, layer1) and the stem max-pool over [256, 64, 112, 112].

    ncu --set full --clock-control none --import-source on -k regex:'bn_|maxpool' --launch-skip 12 -c 12 \
        -o gpurun_out/prof_bn python benchmarks/ncu_bn.py
(the first 12 matching launches are the warm-up pass; never use timings printed under ncu)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from kungfu_b200.models.resnet import BNAct  # noqa: E402
from kungfu_b200.ops.pool import MaxPool3x3s2  # noqa: E402


def one_pass(bn_res, bn_plain, pool, x, res, stem):
    for t in (x, res, stem):
        t.grad = None
    y = bn_res(x, res)            # stats + finalize + apply<res,relu>
    z = bn_plain(y)               # stats + finalize + apply<relu>
    (z.float().mean()).backward()  # bwd_reduce + finalize + bwd_apply for both
    p = pool(stem)
    p.float().mean().backward()
    torch.cuda.synchronize()


def main():
    torch.manual_seed(0)
    dev = "cuda"
    mk = lambda *s: torch.randn(*s, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_()  # noqa: E731
    x, res, stem = mk(256, 256, 56, 56), mk(256, 256, 56, 56), mk(256, 64, 112, 112)
    bn_res, bn_plain = BNAct(256).to(dev, torch.bfloat16).train(), BNAct(256).to(dev, torch.bfloat16).train()
    pool = MaxPool3x3s2()
    one_pass(bn_res, bn_plain, pool, x, res, stem)  # warm-up (skipped by --launch-skip)
    one_pass(bn_res, bn_plain, pool, x, res, stem)
    print("done")


if __name__ == "__main__":
    main()
