"""Process-level runtime singleton: bootstrap, control plane, device plane.

Counterpart of the reference's `kungfu.python` module-level init (srcs/python/kungfu/python/__init__.py:17-28)
+ `kungfu::Peer` (srcs/cpp/include/kungfu/peer.hpp:13-163). A worker can be started three ways:

  * by `kungfu-run` (env contract KUNGFU_SELF_SPEC / KUNGFU_INIT_PEERS / ..., srcs/go/kungfu/env/envs.go:4-18)
    -> native C++ control plane (`_kfrt`), elastic features available;
  * by `torchrun` / `python -m torch.distributed.run` (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)
    -> the same native control plane, peers derived from the env (static membership);
  * standalone (no env) -> single-process mode (reference: env/config.go:25-27,71-80).

The *control plane* carries tiny host messages (barrier, consensus, handle exchange, membership);
the *device plane* (`parallel.mesh`) is peer-mapped HBM driven by our own kernels.
"""
from __future__ import annotations

import atexit
import os
import threading
from typing import Callable, List, Optional

import torch

_lock = threading.RLock()
_runtime: Optional["Runtime"] = None


class Control:
    """Host-side control plane interface."""

    rank = 0
    size = 1
    local_rank = 0
    local_size = 1
    kind = "single"

    def barrier(self) -> None:
        pass

    def all_gather_bytes(self, b: bytes) -> List[bytes]:
        return [bytes(b)]

    def broadcast_bytes(self, b: bytes, root: int = 0) -> bytes:
        return bytes(b)

    def consensus(self, b: bytes) -> bool:
        return True

    def all_reduce_cpu_(self, t: torch.Tensor, op: str = "sum", name: str = "") -> torch.Tensor:
        return t

    def close(self) -> None:
        pass


class TorchDistControl(Control):
    """Control plane over a torch.distributed gloo group (torchrun launches; also the fallback when the
    native runtime is not built)."""

    kind = "torchdist"

    def __init__(self):
        import torch.distributed as dist

        self._dist = dist
        self._own = False
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            dist.init_process_group(backend="gloo")
            self._own = True
            self._group = dist.group.WORLD
        else:
            be = dist.get_backend()
            self._group = dist.group.WORLD if "gloo" in str(be) else dist.new_group(backend="gloo")
        self.rank = dist.get_rank()
        self.size = dist.get_world_size()
        self.local_rank = int(os.environ.get("LOCAL_RANK", self.rank))
        self.local_size = int(os.environ.get("LOCAL_WORLD_SIZE", self.size))

    def barrier(self) -> None:
        self._dist.barrier(group=self._group)

    def all_gather_bytes(self, b: bytes) -> List[bytes]:
        out: List[Optional[bytes]] = [None] * self.size
        self._dist.all_gather_object(out, bytes(b), group=self._group)
        return [bytes(x) for x in out]  # type: ignore[arg-type]

    def broadcast_bytes(self, b: bytes, root: int = 0) -> bytes:
        box = [bytes(b)]
        self._dist.broadcast_object_list(box, src=root, group=self._group)
        return bytes(box[0])

    def consensus(self, b: bytes) -> bool:
        allb = self.all_gather_bytes(b)
        return all(x == allb[0] for x in allb)

    def all_reduce_cpu_(self, t: torch.Tensor, op: str = "sum", name: str = "") -> torch.Tensor:
        ops = {"sum": self._dist.ReduceOp.SUM, "min": self._dist.ReduceOp.MIN, "max": self._dist.ReduceOp.MAX,
               "prod": self._dist.ReduceOp.PRODUCT}
        self._dist.all_reduce(t, op=ops[op], group=self._group)
        return t

    def close(self) -> None:
        if self._own and self._dist.is_initialized():
            try:
                self._dist.destroy_process_group()
            except Exception:
                pass


# Set by `kungfu_b200.cmd.standby`: a warm standby worker runs the training script ahead of time (imports, model construction, moving it
# to the GPU) and is held HERE, at the first thing that needs an identity, until the launcher hands it one.
_activation_gate: Optional[Callable[[], None]] = None


def _make_control() -> Control:
    global _activation_gate
    gate, _activation_gate = _activation_gate, None
    if gate is not None:
        gate()
    env = os.environ
    want = env.get("KUNGFU_CONTROL", "auto").lower()
    launched_by_kungfu = "KUNGFU_SELF_SPEC" in env
    launched_by_torchrun = "RANK" in env and "WORLD_SIZE" in env
    if want in ("auto", "native"):
        try:
            from .peer import NativeControl, native_available

            if native_available() and (launched_by_kungfu or launched_by_torchrun or want == "native"):
                return NativeControl.from_env()
        except ImportError:
            if want == "native":
                raise
    if launched_by_torchrun and int(env["WORLD_SIZE"]) > 1 or want == "torchdist":
        return TorchDistControl()
    return Control()


class Runtime:
    def __init__(self) -> None:
        self._control: Optional[Control] = None
        self._control_lock = threading.Lock()
        self._gpu = None
        self._finalizers: List[Callable[[], None]] = []

    @property
    def control(self) -> Control:
        """The control plane, created at first use (so that what does not need an identity — `cuda_index()`, building a model — can
        run in a standby worker before the launcher has assigned one)."""
        if self._control is None:
            with self._control_lock:
                if self._control is None:
                    self._control = _make_control()
        return self._control

    @control.setter
    def control(self, c: Control) -> None:
        self._control = c

    # ---- identity --------------------------------------------------------------------------------
    @property
    def rank(self) -> int:
        return self.control.rank

    @property
    def size(self) -> int:
        return self.control.size

    @property
    def local_rank(self) -> int:
        return self.control.local_rank

    @property
    def local_size(self) -> int:
        return self.control.local_size

    # ---- device plane ----------------------------------------------------------------------------
    def cuda_index(self) -> int:
        """Logical CUDA device of this worker. Reference: kungfu_get_cuda_index maps
        KUNGFU_CUDA_VISIBLE_DEVICES through CUDA_VISIBLE_DEVICES (srcs/cpp/src/python/cuda.cpp:29-47).
        On a B200 box every worker sees all GPUs (P2P/NVLS need it) and picks device == local rank."""
        env = os.environ
        if "KUNGFU_CUDA_VISIBLE_DEVICES" in env:
            want = env["KUNGFU_CUDA_VISIBLE_DEVICES"].split(",")[0].strip()
            visible = env.get("CUDA_VISIBLE_DEVICES")
            if visible is None:
                return int(want) if want not in ("", "-1") else -1
            ids = [v.strip() for v in visible.split(",")]
            return ids.index(want) if want in ids else -1
        if "LOCAL_RANK" in env:
            return int(env["LOCAL_RANK"])
        n = torch.cuda.device_count() if torch.cuda.is_available() else 0
        return self.local_rank % n if n else -1

    @property
    def gpu(self):
        """Lazily created device plane (`parallel.mesh.GpuPlane`)."""
        if self._gpu is None:
            from .parallel.mesh import GpuPlane

            self._gpu = GpuPlane(self)
        return self._gpu

    def has_gpu(self) -> bool:
        return torch.cuda.is_available() and self.cuda_index() >= 0

    def on_finalize(self, fn: Callable[[], None]) -> None:
        self._finalizers.append(fn)

    def finalize(self) -> None:
        for fn in reversed(self._finalizers):
            try:
                fn()
            except Exception:
                pass
        self._finalizers.clear()
        if self._gpu is not None:
            try:
                self._gpu.close()
            except Exception:
                pass
            self._gpu = None
        if self._control is not None:
            self._control.close()


def get_runtime() -> Runtime:
    global _runtime
    with _lock:
        if _runtime is None:
            _runtime = Runtime()
            atexit.register(_shutdown)
        return _runtime


def reset_runtime() -> None:
    """Tear the singleton down (tests, `launch_multiprocess`, resize re-init)."""
    global _runtime
    with _lock:
        if _runtime is not None:
            _runtime.finalize()
            _runtime = None


def _shutdown() -> None:
    global _runtime
    if _runtime is not None:
        try:
            _runtime.finalize()
        finally:
            _runtime = None
