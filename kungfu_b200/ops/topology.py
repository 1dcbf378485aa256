"""Topology ops: rank / size / peer info, latency probing, minimum spanning tree, neighbour masks, round robin.

Reference: srcs/python/kungfu/tensorflow/ops/topology.py + __init__.py:49-83 and kernels in
srcs/cpp/src/tensorflow/ops/cpu/topology.cpp (`GetPeerLatencies`, `MinimumSpanningTree` = AllGatherTransform + Prim,
`GetNeighbour`, `RoundRobin`).
"""
from __future__ import annotations

from typing import List

import torch

from ..runtime import get_runtime
from . import collective as coll


def rank() -> int:
    return get_runtime().rank


def cluster_size() -> int:
    return get_runtime().size


def peer_info():
    rt = get_runtime()
    return rt.rank, rt.size


def get_peer_latencies(local_step: int = 0) -> torch.Tensor:
    """float32[np]: ping RTT (seconds) to every peer, 0 for self."""
    rt = get_runtime()
    f = getattr(rt.control, "peer_latencies", None)
    lat = f() if f else [0.0] * rt.size
    return torch.tensor(lat, dtype=torch.float32)


def global_minimum_spanning_tree(self_weights: torch.Tensor) -> torch.Tensor:
    """Every peer contributes its row of weights (e.g. latencies); returns the MST edges int32[np-1, 2], identical on
    all peers (gather -> Prim -> identical result everywhere)."""
    from .. import _kfrt  # type: ignore[attr-defined]

    rt = get_runtime()
    n = rt.size
    w = self_weights.detach().to(torch.float64).cpu().contiguous()
    assert w.numel() == n, "need one weight per peer"
    allw = coll.all_gather(w)
    edges = _kfrt.minimum_spanning_tree(allw.view(-1).tolist(), n)
    return torch.tensor(edges, dtype=torch.int32).view(-1, 2)


def minimum_spanning_tree_father(self_weights: torch.Tensor, root: int = 0) -> List[int]:
    """The MST as a father[] array rooted at `root`, ready for `set_tree`."""
    from .. import _kfrt  # type: ignore[attr-defined]

    rt = get_runtime()
    w = self_weights.detach().to(torch.float64).cpu().contiguous()
    allw = coll.all_gather(w)
    return list(_kfrt.mst_forest_array(allw.view(-1).tolist(), rt.size, root))


def get_neighbour_mask(edges: torch.Tensor) -> torch.Tensor:
    """bool[np]: which peers are adjacent to this rank in the edge list."""
    rt = get_runtime()
    mask = torch.zeros(rt.size, dtype=torch.bool)
    for u, v in edges.view(-1, 2).tolist():
        if u == rt.rank:
            mask[v] = True
        if v == rt.rank:
            mask[u] = True
    return mask


class RoundRobin:
    """Successively returns the indices of the True entries of `mask` (reference: `KungfuRoundRobin`)."""

    def __init__(self) -> None:
        self.pos = 0

    def __call__(self, mask: torch.Tensor) -> int:
        idx = [i for i, m in enumerate(mask.tolist()) if m]
        if not idx:
            return -1
        v = idx[self.pos % len(idx)]
        self.pos += 1
        return v


def round_robin(mask: torch.Tensor, _state={"rr": None}) -> int:
    if _state["rr"] is None:
        _state["rr"] = RoundRobin()
    return _state["rr"](mask)
