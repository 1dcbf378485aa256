"""Elastic-aware data sharding.

Reference: srcs/python/kungfu/tensorflow/v1/datasets/adaptor.py — a dataset pipeline whose `skip(offset)`, `batch` and
`shard(np, rank)` parameters are variables, so that after a resize every worker continues from the global sample offset
with the new (np, rank).
"""
from __future__ import annotations

from typing import Iterator, Tuple

import torch

from ..python import current_cluster_size, current_rank


class ElasticDataAdaptor:
    """Deterministic global order; worker r of np takes samples [offset + step*B*np + r*B, ... + B)."""

    def __init__(self, data: torch.Tensor, labels: torch.Tensor, batch_size: int, offset: int = 0) -> None:
        assert len(data) == len(labels)
        self.data, self.labels = data, labels
        self.batch_size = int(batch_size)
        self.offset = int(offset)

    def set_offset(self, trained_samples: int) -> None:
        self.offset = int(trained_samples)

    def next_batch(self) -> Tuple[torch.Tensor, torch.Tensor]:
        n = len(self.data)
        np_, r, b = current_cluster_size(), current_rank(), self.batch_size
        start = (self.offset + r * b) % n
        idx = (torch.arange(b) + start) % n
        self.offset += b * np_
        return self.data[idx], self.labels[idx]

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        while True:
            yield self.next_batch()
