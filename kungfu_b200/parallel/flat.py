"""Flat gradient / parameter arenas and the bucket plan.

Reference: gradient `fuse`/`defuse` (srcs/python/kungfu/tensorflow/ops/__init__.py:29-46, concat + slice copies
every step) and the per-tensor named all-reduces of the torch path (torch/optimizers/sync_sgd.py:12-22).
Here gradients and parameters are *views* into pre-allocated flat regions of the symmetric arena, laid out in
reverse registration order (≈ the order backward produces them) and cut into buckets; nothing is copied.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import torch

ALIGN_BYTES = 128  # every tensor starts on a 128-byte boundary (cuDNN/cuBLAS fast paths need >= 16)


@dataclass
class Entry:
    param: torch.nn.Parameter
    name: str
    offset: int  # element offset inside the flat region
    numel: int
    bucket: int = -1


@dataclass
class Bucket:
    index: int
    start: int  # element offset (aligned)
    numel: int  # padded element count (multiple of 16 bytes)
    entries: List[int] = field(default_factory=list)

    def nbytes(self, esize: int) -> int:
        return self.numel * esize


class FlatLayout:
    """Layout of one homogeneous (dtype, hyper-parameter group) set of tensors."""

    def __init__(self, named_params: Sequence, dtype: torch.dtype, bucket_bytes: int, reverse: bool = True, tail_bytes: int = 0) -> None:
        self.dtype = dtype
        self.esize = torch.empty((), dtype=dtype).element_size()
        align = ALIGN_BYTES // self.esize
        items = list(named_params)
        if reverse:
            items = items[::-1]
        self.entries: List[Entry] = []
        self.buckets: List[Bucket] = []
        cur = 0
        bstart = 0
        bucket_elems = max(align, bucket_bytes // self.esize)
        cur_bucket = Bucket(0, 0, 0)
        for name, p in items:
            off = (cur + align - 1) // align * align
            e = Entry(p, name, off, p.numel(), cur_bucket.index)
            self.entries.append(e)
            cur_bucket.entries.append(len(self.entries) - 1)
            cur = off + p.numel()
            if cur - bstart >= bucket_elems:
                end = (cur + align - 1) // align * align
                cur_bucket.start, cur_bucket.numel = bstart, end - bstart
                self.buckets.append(cur_bucket)
                cur = bstart = end
                cur_bucket = Bucket(len(self.buckets), bstart, 0)
        if cur_bucket.entries:
            end = (cur + align - 1) // align * align
            cur_bucket.start, cur_bucket.numel = bstart, end - bstart
            self.buckets.append(cur_bucket)
            cur = end
        self.total = cur  # elements
        # The LAST bucket is the one backward produces last (the first layers of the model): nothing is left to hide its
        # collective under, so its size is exposed step time. Cut a small tail off it: the bulk of the remainder still overlaps
        # with the end of backward, only `tail_bytes` worth of gradients is reduced after the last backward kernel.
        if tail_bytes > 0 and len(self.buckets) >= 1:
            last = self.buckets[-1]
            tail_elems = max(align, tail_bytes // self.esize)
            if last.numel > 2 * tail_elems and len(last.entries) > 1:
                keep = list(last.entries)
                moved = []
                size = 0
                while len(keep) > 1 and size + self.entries[keep[-1]].numel <= tail_elems:
                    size += self.entries[keep[-1]].numel
                    moved.insert(0, keep.pop())
                if moved:
                    split = self.entries[moved[0]].offset  # aligned: every entry starts on an `align` boundary
                    nb = Bucket(len(self.buckets), split, last.start + last.numel - split, moved)
                    last.entries, last.numel = keep, split - last.start
                    for ei in moved:
                        self.entries[ei].bucket = nb.index
                    self.buckets.append(nb)

    @property
    def nbytes(self) -> int:
        return self.total * self.esize

    def digest(self) -> bytes:
        """Order/shape digest; all ranks must agree on it (consensus-checked once, replacing the
        reference's per-step NCCL order negotiation, srcs/cpp/src/nccl/scheduler.cpp:93-119)."""
        import hashlib

        h = hashlib.sha1()
        h.update(str(self.dtype).encode())
        for e in self.entries:
            h.update(f"{e.name}:{e.offset}:{e.numel}:{e.bucket};".encode())
        for b in self.buckets:
            h.update(f"B{b.index}:{b.start}:{b.numel};".encode())
        return h.digest()

    def shard_elems(self, bucket: Bucket, nranks: int) -> int:
        """fp32 state elements a rank owns for `bucket` (ceil(nvec / nranks) vectors)."""
        lanes = 16 // self.esize
        nvec = bucket.numel // lanes
        per = (nvec + nranks - 1) // nranks
        return per * lanes

    def shard_range(self, bucket: Bucket, nranks: int, rank: int):
        """Element range [lo, hi) of `bucket` (relative to the flat region) owned by `rank`."""
        lanes = 16 // self.esize
        nvec = bucket.numel // lanes
        per = (nvec + nranks - 1) // nranks
        s0 = min(rank * per, nvec)
        s1 = min(s0 + per, nvec)
        return bucket.start + s0 * lanes, bucket.start + s1 * lanes


def group_params(param_groups, named_parameters) -> List[Dict]:
    """Split optimizer param_groups into homogeneous (group, dtype) sets with stable names."""
    names = {id(p): n for n, p in named_parameters} if named_parameters is not None else {}
    out = []
    for gi, g in enumerate(param_groups):
        by_dtype: Dict[torch.dtype, List] = {}
        for pi, p in enumerate(g["params"]):
            if not p.requires_grad:
                continue
            by_dtype.setdefault(p.dtype, []).append((names.get(id(p), f"group{gi}.param{pi}"), p))
        for dt, items in by_dtype.items():
            out.append({"group": gi, "dtype": dt, "items": items})
    return out
