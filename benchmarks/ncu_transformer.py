"""Tiny workload for an Nsight Compute capture of the fused transformer kernels at the BERT-large shapes (batch 32 x seq 128 = 4096 tokens,
H = 1024, I = 4096, 16 heads): one forward + backward of each op, cold L2 between ops (a 256 MiB fill).

    ncu --set full --clock-control none --import-source on -k regex:'add_ln_|bias_gelu_|qkv_pack|colsum_finalize' -c 12 -f \
        -o gpurun_out/ncu_transformer python benchmarks/ncu_transformer.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kungfu_b200.ops.transformer import add_layernorm, bias_gelu, qkv_projection  # noqa: E402

dev = "cuda"
B, S, H, I, heads = 32, 128, 1024, 4096, 16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def cold():
    flush.fill_(1)
    torch.cuda.synchronize()


def t(*shape, grad=True):
    return (torch.randn(*shape, device=dev) * 0.5).bfloat16().requires_grad_(grad)


x, r, g, b, lb = t(B * S, H), t(B * S, H), t(H), t(H), t(H)
cold()
y = add_layernorm(x, r, g, b, 1e-12, lin_bias=lb)
dy = torch.randn_like(y)
cold()
y.backward(dy)
xi, bi = t(B * S, I), t(I)
cold()
yi = bias_gelu(xi, bi)
dyi = torch.randn_like(yi)
cold()
yi.backward(dyi)
xq, wq, bq = t(B, S, H), t(3 * H, H), t(3 * H)
q, k, v = qkv_projection(xq, wq, bq, heads)
a = torch.nn.functional.scaled_dot_product_attention(q, k, v)
da = torch.randn_like(a)
cold()
a.backward(da)
torch.cuda.synchronize()
print("ok", float(x.grad.float().abs().mean()), float(xi.grad.float().abs().mean()), float(xq.grad.float().abs().mean()))
