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
    if r != 0 and not all_ranks:
        return None
    if all_ranks:
        path = f"{path}.rank{r}"
    blob = {
        "model": {k: v.detach().cpu() for k, v in model.state_dict().items()},
        "optimizer": optimizer.state_dict() if optimizer is not None and getattr(optimizer, "engine", None) is None else None,
        "step": int(step), "epoch": int(epoch), "trained_samples": int(trained_samples), "extra": extra or {},
    }
    eng = getattr(optimizer, "engine", None) if optimizer is not None else None
    if eng is not None:
        blob["engine_state"] = {k: v.detach().cpu() for k, v in eng.gather_state().items()}
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
        if optimizer is not None and blob.get("optimizer") is not None and getattr(optimizer, "engine", None) is None:
            optimizer.load_state_dict(blob["optimizer"])
        meta.update({k: blob.get(k, meta[k]) for k in ("step", "epoch", "trained_samples", "extra")})
        meta["found"] = True
    if broadcast:
        BroadcastGlobalVariablesOp(model, optimizer if getattr(optimizer, "engine", None) is None else None, 0)
        run_barrier()
    return meta
