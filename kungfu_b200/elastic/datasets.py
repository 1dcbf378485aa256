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


class BaseDatasetAdaptor:
    """Reference-shaped adaptor (v1/datasets/adaptor.py:4-34): batch size, shard count, shard id and the global sample offset are
    mutable state; `__call__(dataset)` yields this worker's batches of an indexable dataset from the current offset on, and the
    `create_*` methods return the callables the reference builds as TF ops."""

    def __init__(self, batch_size: int = 10, shard_count: int = 1, shard_id: int = 0) -> None:
        self._batch_size, self._shard_count, self._shard_id, self._offset = int(batch_size), int(shard_count), int(shard_id), 0

    def create_update_offset(self):
        def op() -> int:
            self._offset += self._batch_size * self._shard_count
            return self._offset

        return op

    def create_rewind(self):
        def op() -> int:
            self._offset = 0
            return 0

        return op

    def create_update_topology(self):
        raise RuntimeError("Not implemented")  # subclasses decide how (np, rank, batch size) follow the cluster

    def debug(self) -> str:
        s = "offset=%d, np=%d, rank=%d, batch size=%d" % (self._offset, self._shard_count, self._shard_id, self._batch_size)
        print(s)
        return s

    def __call__(self, ds):
        """ds.skip(offset).batch(batch_size).shard(np, rank): batch k of the remaining data belongs to shard k % np."""
        n = len(ds)

        def gen():
            k = 0
            start = self._offset
            while start + self._batch_size <= n:
                if k % self._shard_count == self._shard_id:
                    yield ds[start:start + self._batch_size]
                start += self._batch_size
                k += 1

        return gen()


class ExampleDatasetAdaptor(BaseDatasetAdaptor):
    """The reference's demonstration subclass (adaptor.py:36-47): one more worker, batch size + 2, shard id rotated."""

    def create_update_topology(self):
        def op() -> None:
            self._shard_count += 1
            self._shard_id = (self._shard_id + 1) % self._shard_count
            self._batch_size += 2

        return op


class ClusterDatasetAdaptor(BaseDatasetAdaptor):
    """What a real job wants: (np, rank) follow the live cluster after every resize."""

    def create_update_topology(self):
        def op() -> None:
            self._shard_count, self._shard_id = current_cluster_size(), current_rank()

        return op
