"""Named global variables: the channel through which policies, monitors and the training loop exchange scalars
(batch size, gradient noise scale, trained samples).

Reference: srcs/python/kungfu/tensorflow/variables.py:1-123 keeps them in TF graph collections (`get_global_variable`,
`create_global_variable`, `get_or_create_global_variable`, `eval_global_variable`, `GraphKeys`, `get_or_create_batch_size`,
`eval_batch_size`, `eval_gradient_noise_scale`, `create_setter`, `create_assign_op_for`, `create_placeholder_for`). Without a
graph, the collection is a process-wide registry of small tensors; "eval" reads the value, a "setter" writes it in place
(so that anything holding the tensor — a CUDA graph included — sees the update), and `sync_global_variables` makes a
variable identical on all workers (the reference's variables live in `GLOBAL_VARIABLES` and ride along with
`BroadcastGlobalVariablesOp`).
"""
from __future__ import annotations

import threading
from typing import Callable, Dict, Optional, Tuple

import torch

_lock = threading.RLock()
_registry: Dict[str, torch.Tensor] = {}


class GraphKeys:
    BATCH_SIZE = "kungfu_batch_size"
    GRADIENT_NOISE_SCALE = "kungfu_gradient_noise_scale"
    TOTAL_SAMPLES = "kungfu_total_samples"
    TRAINED_SAMPLES = "kungfu_trained_samples"


def get_global_variable(name: str) -> Optional[torch.Tensor]:
    """The tensor registered under `name`, or None."""
    with _lock:
        return _registry.get(name)


def create_global_variable(name: str, shape=(), dtype: torch.dtype = torch.float32, init=None, device="cpu") -> torch.Tensor:
    """Register a new variable; `init` is a number or a tensor (then `shape` is taken from it, like a constant initializer)."""
    with _lock:
        if name in _registry:
            raise ValueError('"%s" already exists.' % name)
        if init is None:
            v = torch.zeros(tuple(shape), dtype=dtype, device=device)
        else:
            t = torch.as_tensor(init, dtype=dtype)
            v = t.detach().clone().to(device) if t.dim() or not tuple(shape) else torch.full(tuple(shape), float(init), dtype=dtype, device=device)
        v.requires_grad_(False)
        _registry[name] = v
        return v


def get_or_create_global_variable(name: str, shape=(), dtype: torch.dtype = torch.float32, init=None, device="cpu") -> torch.Tensor:
    """Existing variable `name`, or a new one created with the given shape / dtype / initial value."""
    with _lock:
        v = _registry.get(name)
        return v if v is not None else create_global_variable(name, shape, dtype, init, device)


def eval_global_variable(name: str):
    """Current value: a Python scalar for one-element variables, a CPU copy otherwise; raises when the variable does not exist."""
    v = get_global_variable(name)
    if v is None:
        raise RuntimeError('"%s" not exist' % name)
    return v.item() if v.numel() == 1 else v.detach().cpu().clone()


def bind_global_variable(name: str, tensor: torch.Tensor) -> torch.Tensor:
    """Make `name` refer to an existing tensor (e.g. a device scalar that a kernel updates); replaces a previous binding."""
    with _lock:
        _registry[name] = tensor
        return tensor


def remove_global_variable(name: str) -> None:
    """Forget `name` (tests; a new job in the same process)."""
    with _lock:
        _registry.pop(name, None)


def global_variables() -> Dict[str, torch.Tensor]:
    """Snapshot of the registry {name: tensor}."""
    with _lock:
        return dict(_registry)


# ---- assignment helpers (reference: create_placeholder_for / create_assign_op_for / create_setter) ----------------------
def create_placeholder_for(t: torch.Tensor) -> torch.Tensor:
    """A staging tensor shaped and typed like `t` (pinned when `t` lives on a GPU, so that the assign is an async H2D copy)."""
    p = torch.empty(t.shape, dtype=t.dtype)
    return p.pin_memory() if t.is_cuda else p


def create_assign_op_for(t: torch.Tensor) -> Tuple[Callable[[], None], torch.Tensor]:
    """(op, placeholder): fill the placeholder, call op() to copy it into `t` (async H2D for device variables)."""
    p = create_placeholder_for(t)

    def op() -> None:
        t.copy_(p, non_blocking=True)

    return op, p


def create_setter(v: torch.Tensor) -> Callable:
    """`set_value(value)` that writes `value` into the variable in place."""
    op, place = create_assign_op_for(v)

    def set_value(value) -> None:
        place.copy_(torch.as_tensor(value, dtype=v.dtype).reshape(v.shape))
        op()

    return set_value


# ---- well-known variables -------------------------------------------------------------------------------------------------
def get_or_create_batch_size(init=None) -> torch.Tensor:
    """The int32 scalar `kungfu_batch_size` (created on first use)."""
    return get_or_create_global_variable(GraphKeys.BATCH_SIZE, shape=(), dtype=torch.int32, init=init)


def get_batch_size() -> Optional[torch.Tensor]:
    """The batch size variable (None before anybody created it)."""
    return get_global_variable(GraphKeys.BATCH_SIZE)


def eval_batch_size() -> int:
    """Value of `kungfu_batch_size`."""
    return int(eval_global_variable(GraphKeys.BATCH_SIZE))


def eval_gradient_noise_scale() -> float:
    """Value of `kungfu_gradient_noise_scale` (synchronises when the estimate lives on the GPU)."""
    return float(eval_global_variable(GraphKeys.GRADIENT_NOISE_SCALE))


def sync_global_variables(root: int = 0, names=None) -> None:
    """Broadcast the registry (or `names`) from `root`; newcomers of an elastic job call this after `sync_after_resize`."""
    from .ops import collective as coll

    with _lock:
        items = sorted((n, v) for n, v in _registry.items() if names is None or n in names)
    for n, v in items:
        coll.broadcast_(v, root, "kungfu::var::" + n)
