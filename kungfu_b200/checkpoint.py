"""Checkpoint / resume helpers.

The reference runtime has no checkpoint format: examples save per-rank model files each epoch and
reload on `--restart 1` (examples/Failure_recovery_examples/eager.py:76-104). Here: rank 0 writes one atomic file, every
rank can load it and the state is re-synchronised by broadcast — used by the failure-recovery loop and the elastic tests.
"""
from __future__ import annotations

import os
import tempfile
from typing import Any, Dict, Optional

import torch

from .initializer import BroadcastGlobalVariablesOp
from .python import current_rank, run_barrier


def save(path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, step: int = 0, epoch: int = 0,
         trained_samples: int = 0, extra: Optional[Dict[str, Any]] = None, all_ranks: bool = False) -> Optional[str]:
    """Rank 0 (or every rank with `all_ranks`, suffixing `.rank<r>`) writes {model, optimizer, progress}."""
    r = current_rank()
    eng = getattr(optimizer, "engine", None) if optimizer is not None else None
    # COLLECTIVE part first, on every rank: a fused engine shards fp32 master weights / optimizer state across ranks and
    # un-shards them with device collectives — rank 0 must not be alone in there
    engine_state = eng.state_dict() if eng is not None else None
    if r != 0 and not all_ranks:
        return None
    if all_ranks:
        path = f"{path}.rank{r}"
    # the torch optimizer's own state is what counts whenever the engine does not hold it (no engine, or 'grads' mode where the
    # engine only reduces and the wrapped optimizer steps)
    inner = optimizer is not None and (eng is None or eng.mode == "grads")
    blob = {
        "model": {k: v.detach().cpu() for k, v in model.state_dict().items()},
        "optimizer": optimizer.state_dict() if inner else None,
        "step": int(step), "epoch": int(epoch), "trained_samples": int(trained_samples), "extra": extra or {},
    }
    if engine_state is not None:
        blob["engine_state"] = engine_state
    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    fd, tmp = tempfile.mkstemp(dir=d, suffix=".tmp")
    os.close(fd)
    torch.save(blob, tmp)
    os.replace(tmp, path)  # atomic: a crash never leaves a torn checkpoint
    return path


def load(path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, broadcast: bool = True,
         map_location="cpu") -> Dict[str, Any]:
    """Load (rank-local read) then make rank 0's state authoritative by broadcast. Missing file -> progress zeros."""
    meta = {"step": 0, "epoch": 0, "trained_samples": 0, "extra": {}, "found": False}
    if os.path.exists(path):
        blob = torch.load(path, map_location=map_location, weights_only=False)
        sd = model.state_dict()
        for k, v in blob["model"].items():
            if k in sd:
                sd[k].copy_(v.to(sd[k].device, sd[k].dtype))
        eng = getattr(optimizer, "engine", None) if optimizer is not None else None
        if optimizer is not None and blob.get("optimizer") is not None and (eng is None or eng.mode == "grads"):
            optimizer.load_state_dict(blob["optimizer"])
        if eng is not None and blob.get("engine_state") is not None and eng.mode != "grads":
            eng.load_state_dict(blob["engine_state"])  # fp32 master shards, momentum / Adam moments, device step counters
        meta.update({k: blob.get(k, meta[k]) for k in ("step", "epoch", "trained_samples", "extra")})
        meta["found"] = True
    if broadcast:
        # every rank read the same file, so the engine state needs no broadcast; model tensors are re-broadcast so that a rank
        # that found no (or an older) file still starts from rank 0's weights
        BroadcastGlobalVariablesOp(model, optimizer if getattr(optimizer, "engine", None) is None else None, 0)
        run_barrier()
    return meta
