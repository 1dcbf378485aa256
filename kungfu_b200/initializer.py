"""Initial state synchronisation.

Reference: `kungfu.tensorflow.initializer` (srcs/python/kungfu/tensorflow/initializer/__init__.py:13-99):
`BroadcastGlobalVariablesOp`, `broadcast_variables`, `BroadcastGlobalVariablesHook`, `BroadcastGlobalVariablesCallback`
(fires after the first batch so optimizer slots exist). Torch: `broadcast_parameters(state_dict)`
(srcs/python/kungfu/torch/ops/collective.py:42-47).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from .ops import collective as coll


def broadcast_variables(tensors: Iterable[torch.Tensor], root: int = 0) -> None:
    """Broadcast a list of tensors in place from `root` (reference: initializer `broadcast_variables`)."""
    for t in tensors:
        coll.broadcast_(t, root)


def broadcast_parameters(state_dict, root: int = 0) -> None:
    """Broadcast every tensor of a state dict (parameters and buffers) from `root`."""
    for _, v in state_dict.items():
        if isinstance(v, torch.Tensor):
            coll.broadcast_(v, root)


def _numeric_hyper(group) -> list:
    """(key, index-or-None) of every numeric hyper-parameter of a param group, in a rank-independent order."""
    slots = []
    for k in sorted(group.keys()):
        v = group[k]
        if k == "params" or isinstance(v, bool) or v is None:
            continue
        if isinstance(v, (int, float)):
            slots.append((k, None))
        elif isinstance(v, (tuple, list)) and v and all(isinstance(x, (int, float)) and not isinstance(x, bool) for x in v):
            slots += [(k, i) for i in range(len(v))]
    return slots


def broadcast_optimizer_state(optimizer: torch.optim.Optimizer, root: int = 0) -> None:
    """Broadcast every tensor in the optimizer state (momentum buffers, Adam moments, step counters) and the numeric
    hyper-parameters of the param groups (lr, momentum, betas, weight decay, ...): in the reference these are graph variables
    and `BroadcastGlobalVariablesOp` covers them; a learning rate that a scheduler already moved on the root must reach
    workers that join or restart.

    torch creates optimizer state lazily, at the first `step()`: a worker that joins a running job has none while the root has
    momentum for every parameter. The root therefore announces the LAYOUT of its state first (which parameter has which slot, shape,
    dtype, device kind); everybody else creates what it lacks, and only then are the tensors broadcast — all members issue the same
    collectives in the same order whatever they held before."""
    import json

    from .runtime import get_runtime

    rt = get_runtime()
    layout = []
    for gi, group in enumerate(optimizer.param_groups):
        for pi, p in enumerate(group["params"]):
            st = optimizer.state.get(p, {})
            for k in sorted(st.keys()):
                v = st[k]
                if isinstance(v, torch.Tensor):
                    layout.append([gi, pi, k, "t", list(v.shape), str(v.dtype).replace("torch.", ""), v.device.type])
                elif isinstance(v, (int, float)) and not isinstance(v, bool):
                    layout.append([gi, pi, k, "s", type(v).__name__, v])
    if rt.size > 1:
        layout = json.loads(rt.control.broadcast_bytes(json.dumps(layout).encode(), root).decode())
    for e in layout:
        gi, pi, k, kind = e[:4]
        p = optimizer.param_groups[gi]["params"][pi]
        st = optimizer.state[p]
        if kind == "s":
            st[k] = int(e[5]) if e[4] == "int" else float(e[5])
            continue
        shape, dtype, devkind = e[4], getattr(torch, e[5]), e[6]
        v = st.get(k)
        dev = p.device if devkind == p.device.type else torch.device(devkind)
        if not isinstance(v, torch.Tensor) or list(v.shape) != shape or v.dtype != dtype or v.device.type != dev.type:
            v = torch.zeros(shape, dtype=dtype, device=dev)
            st[k] = v
        coll.broadcast_(v, root)
    for group in optimizer.param_groups:
        slots = _numeric_hyper(group)
        if not slots:
            continue
        vals = torch.tensor([float(group[k] if i is None else group[k][i]) for k, i in slots], dtype=torch.float64)
        coll.broadcast_(vals, root)
        for (k, i), v in zip(slots, vals.tolist()):
            if i is None:
                group[k] = type(group[k])(v)
            else:
                seq = list(group[k])
                seq[i] = type(seq[i])(v)
                group[k] = type(group[k])(seq)


def BroadcastGlobalVariablesOp(model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, root: int = 0) -> None:
    """Make every worker start from rank `root`'s model (parameters AND buffers) and optimizer state."""
    broadcast_parameters(model.state_dict(), root)
    if optimizer is not None:
        broadcast_optimizer_state(optimizer, root)


class BroadcastGlobalVariablesCallback:
    """Training-loop callback: broadcasts once, after the first batch (so lazily created optimizer slots exist)."""

    def __init__(self, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, root: int = 0) -> None:
        self.model, self.optimizer, self.root = model, optimizer, root
        self.done = False

    def on_batch_end(self, batch: int = 0) -> None:
        if not self.done:
            BroadcastGlobalVariablesOp(self.model, self.optimizer, self.root)
            self.done = True


BroadcastGlobalVariablesHook = BroadcastGlobalVariablesCallback
