"""BERT encoder (base / large) for the PairAveraging configuration of BASELINE.json ("BERT-large PairAveragingOptimizer
async gossip"). The reference ships no BERT model — only a gradient-size list of BERT-base + Adam slots
(srcs/python/kungfu/tensorflow/v1/benchmarks/model_sizes.py:27-81, tests/cpp/integration/bert.hpp) and a link to an external
example (README.md:136-137); BERT-large shapes therefore come from this module: 24 layers, hidden 1024, 16 heads,
≈335 M parameters. Attention uses `F.scaled_dot_product_attention` (library flash kernels), the projections are plain library
GEMMs; everything between them runs on our fused kernels (`ops/transformer.py`: residual + LayerNorm in one pass, FFN bias +
GELU in one pass, both with their parameter gradients produced inside the backward pass), and the gossip step over the
parameter set is our TMA pull + `pair_avg_kernel`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class BertConfig:
    vocab_size: int = 30522
    hidden: int = 1024
    layers: int = 24
    heads: int = 16
    intermediate: int = 4096
    max_position: int = 512
    type_vocab: int = 2
    num_labels: int = 2
    dropout: float = 0.0
    fused_mlp: bool = False  # FFN up-projection + GELU on the tcgen05 GEMM with a fused epilogue (ops/linear_act.py)
    fused_ops: bool = True   # residual+LayerNorm and bias+GELU on our kernels (fall back to PyTorch off-GPU / in fp32)

    @staticmethod
    def large() -> "BertConfig":
        return BertConfig()

    @staticmethod
    def base() -> "BertConfig":
        return BertConfig(hidden=768, layers=12, heads=12, intermediate=3072)

    @staticmethod
    def tiny() -> "BertConfig":
        return BertConfig(vocab_size=1000, hidden=64, layers=2, heads=4, intermediate=128, max_position=64)


class BertLayer(nn.Module):
    def __init__(self, c: BertConfig) -> None:
        super().__init__()
        self.heads = c.heads
        self.qkv = nn.Linear(c.hidden, 3 * c.hidden)
        self.out = nn.Linear(c.hidden, c.hidden)
        self.ln1 = nn.LayerNorm(c.hidden, eps=1e-12)
        self.fused_mlp = c.fused_mlp
        if c.fused_mlp:
            from ..ops.linear_act import FusedLinear

            self.ff1 = FusedLinear(c.hidden, c.intermediate, act="gelu")
        else:
            self.ff1 = nn.Linear(c.hidden, c.intermediate)
        self.ff2 = nn.Linear(c.intermediate, c.hidden)
        self.ln2 = nn.LayerNorm(c.hidden, eps=1e-12)

        self.fused_ops = c.fused_ops

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, S, H = x.shape
        if not self.fused_ops:
            q, k, v = self.qkv(x).view(B, S, 3, self.heads, H // self.heads).permute(2, 0, 3, 1, 4)
            a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, H)
            x = self.ln1(x + self.out(a))
            h = self.ff1(x) if self.fused_mlp else F.gelu(self.ff1(x))
            return self.ln2(x + self.ff2(h))
        from ..ops.transformer import add_layernorm, bias_gelu, qkv_projection

        # the biases of the two output GEMMs ride in the residual+LayerNorm kernels, the gradient of the fused QKV projection is packed
        # by one kernel (no cat / layout copy / column-sum kernels in backward)
        q, k, v = qkv_projection(x, self.qkv.weight, self.qkv.bias, self.heads)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, H)
        x = add_layernorm(F.linear(a, self.out.weight), x, self.ln1.weight, self.ln1.bias, self.ln1.eps, lin_bias=self.out.bias)
        h = self.ff1(x) if self.fused_mlp else bias_gelu(F.linear(x, self.ff1.weight), self.ff1.bias)
        return add_layernorm(F.linear(h, self.ff2.weight), x, self.ln2.weight, self.ln2.bias, self.ln2.eps, lin_bias=self.ff2.bias)


class Bert(nn.Module):
    def __init__(self, c: BertConfig) -> None:
        super().__init__()
        self.cfg = c
        self.word = nn.Embedding(c.vocab_size, c.hidden)
        self.pos = nn.Embedding(c.max_position, c.hidden)
        self.typ = nn.Embedding(c.type_vocab, c.hidden)
        self.ln = nn.LayerNorm(c.hidden, eps=1e-12)
        self.layers = nn.ModuleList([BertLayer(c) for _ in range(c.layers)])
        self.pooler = nn.Linear(c.hidden, c.hidden)
        self.classifier = nn.Linear(c.hidden, c.num_labels)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        B, S = ids.shape
        pos = torch.arange(S, device=ids.device)
        emb = self.word(ids) + self.pos(pos)[None] + self.typ.weight[0][None, None]
        if self.cfg.fused_ops:
            from ..ops.transformer import add_layernorm

            x = add_layernorm(emb, None, self.ln.weight, self.ln.bias, self.ln.eps)
        else:
            x = self.ln(emb)
        for l in self.layers:
            x = l(x)
        return self.classifier(torch.tanh(self.pooler(x[:, 0])))


def bert_large() -> Bert:
    return Bert(BertConfig.large())


def bert_base() -> Bert:
    return Bert(BertConfig.base())


def parameter_sizes(c: BertConfig) -> List[int]:
    """Parameter tensor sizes without instantiating the weights."""
    H, I = c.hidden, c.intermediate
    s = [c.vocab_size * H, c.max_position * H, c.type_vocab * H, H, H]
    for _ in range(c.layers):
        s += [3 * H * H, 3 * H, H * H, H, H, H, H * I, I, I * H, H, H, H]
    s += [H * H, H, H * c.num_labels, c.num_labels]
    return s
