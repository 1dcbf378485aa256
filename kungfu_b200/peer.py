"""Python face of the native CPU runtime (`_kfrt.Peer`): the control plane used under `kungfu-run`, under torchrun
and in GPU-less tests.

Reference: `kungfu.python` ctypes layer over `libkungfu` / `kungfu::Peer` (srcs/python/kungfu/python/__init__.py:17-155,
srcs/cpp/src/python/init.cpp:30-34).
"""
from __future__ import annotations

import itertools
import os
from typing import List, Optional, Tuple

import torch

from .runtime import Control
from .utils import dtypes

try:
    from . import _kfrt  # type: ignore[attr-defined]
except ImportError:  # pragma: no cover - build missing
    _kfrt = None


def native_available() -> bool:
    return _kfrt is not None


_OPS = {"sum": dtypes.SUM, "min": dtypes.MIN, "max": dtypes.MAX, "prod": dtypes.PROD}
_seq = {"version": None, "counts": {}}


def _auto_name(prefix: str, name: str, version: int = 0) -> str:
    """Unique op names: the rendezvous key is (peer, name), so two in-flight ops must never share one
    (reference: Appendix B — concurrency safety comes from unique tensor names). Numbering restarts with every
    cluster version so that survivors and late joiners of an elastic job generate the same names, and it is kept PER KIND
    of operation: ops that only some ranks take part in (cross-host / subset all-reduce on the hosts' first ranks, local
    reduce) must not shift the numbers of the ops everybody calls."""
    if name:
        return name
    if _seq["version"] != version:
        _seq["version"] = version
        _seq["counts"] = {}
    k = _seq["counts"].get(prefix, 0)
    _seq["counts"][prefix] = k + 1
    return f"{prefix}::v{version}::{k}"


class NativeControl(Control):
    kind = "native"

    def __init__(self, peer) -> None:
        self.peer = peer
        self._version = 0
        self.peer.start()
        self._refresh()
        try:
            _kfrt.apply_affinity_from_env(self.local_rank, self.local_size)
        except Exception:
            pass

    # ---- construction ---------------------------------------------------------------------------------
    @classmethod
    def from_env(cls) -> "NativeControl":
        return cls(_kfrt.Peer.from_env())

    @classmethod
    def single_machine(cls, rank: int, size: int, base_port: int = 0) -> "NativeControl":
        return cls(_kfrt.Peer.single_machine(rank, size, base_port))

    def _refresh(self) -> None:
        if self.peer.detached():
            return
        self.rank = self.peer.rank()
        self.size = self.peer.size()
        self.local_rank = self.peer.local_rank()
        self.local_size = self.peer.local_size()
        self._version = self.peer.cluster_version()

    # ---- identity -------------------------------------------------------------------------------------
    def uid(self) -> int:
        return self.peer.uid()

    def detached(self) -> bool:
        return self.peer.detached()

    def host_count(self) -> int:
        return self.peer.host_count()

    def cluster_version(self) -> int:
        return self.peer.cluster_version()

    def launcher(self) -> str:
        return self.peer.launcher()

    # ---- collectives on host memory -------------------------------------------------------------------
    def barrier(self) -> None:
        self.peer.barrier()

    def all_gather_bytes(self, b: bytes) -> List[bytes]:
        return [bytes(x) for x in self.peer.all_gather_bytes(bytes(b), "")]

    def broadcast_bytes(self, b: bytes, root: int = 0) -> bytes:
        return self.all_gather_bytes(b if self.rank == root else b"")[root]

    def consensus(self, b: bytes) -> bool:
        return bool(self.peer.consensus(bytes(b), ""))

    def all_reduce_cpu_(self, t: torch.Tensor, op: str = "sum", name: str = "") -> torch.Tensor:
        assert not t.is_cuda and t.is_contiguous()
        self.peer.all_reduce(t.data_ptr(), t.data_ptr(), t.numel(), dtypes.dtype_code(t.dtype), _OPS[op], _auto_name("allreduce", name, self._version))
        return t

    def all_reduce_cpu_async_(self, t: torch.Tensor, op: str = "sum", name: str = "") -> int:
        assert not t.is_cuda and t.is_contiguous()
        return self.peer.all_reduce_async(t.data_ptr(), t.data_ptr(), t.numel(), dtypes.dtype_code(t.dtype), _OPS[op],
                                          _auto_name("allreduce", name, self._version))

    def monitored_all_reduce_cpu_(self, t: torch.Tensor, tree: Optional[List[int]] = None, op: str = "sum", name: str = "") -> torch.Tensor:
        self.peer.monitored_all_reduce(t.data_ptr(), t.data_ptr(), t.numel(), dtypes.dtype_code(t.dtype), _OPS[op],
                                       _auto_name("mallreduce", name, self._version), list(tree or []))
        return t

    def subset_all_reduce_cpu_(self, t: torch.Tensor, forest: List[int], op: str = "sum", name: str = "") -> torch.Tensor:
        """All-reduce inside every tree of `forest` (father array, father[i] == i for roots). Only members of multi-node
        trees need to call it."""
        assert not t.is_cuda and t.is_contiguous()
        self.peer.subset_all_reduce(t.data_ptr(), t.data_ptr(), t.numel(), dtypes.dtype_code(t.dtype), _OPS[op],
                                    _auto_name("sallreduce", name, self._version), [int(x) for x in forest])
        return t

    def cross_all_reduce_cpu_(self, t: torch.Tensor, op: str = "sum", name: str = "") -> torch.Tensor:
        self.peer.cross_all_reduce(t.data_ptr(), t.data_ptr(), t.numel(), dtypes.dtype_code(t.dtype), _OPS[op], _auto_name("xallreduce", name, self._version))
        return t

    def reduce_cpu(self, t: torch.Tensor, op: str = "sum", name: str = "") -> torch.Tensor:
        out = torch.empty_like(t)
        self.peer.reduce(t.data_ptr(), out.data_ptr(), t.numel(), dtypes.dtype_code(t.dtype), _OPS[op], _auto_name("reduce", name, self._version))
        return out

    def broadcast_cpu_(self, t: torch.Tensor, name: str = "") -> torch.Tensor:
        self.peer.broadcast(t.data_ptr(), t.data_ptr(), t.numel(), dtypes.dtype_code(t.dtype), _auto_name("bcast", name, self._version))
        return t

    def local_reduce_cpu(self, t: torch.Tensor, op: str = "sum", name: str = "") -> torch.Tensor:
        out = t.clone()
        self.peer.local_reduce(t.data_ptr(), out.data_ptr(), t.numel(), dtypes.dtype_code(t.dtype), _OPS[op], _auto_name("lreduce", name, self._version))
        return out

    def local_broadcast_cpu_(self, t: torch.Tensor, name: str = "") -> torch.Tensor:
        self.peer.local_broadcast(t.data_ptr(), t.data_ptr(), t.numel(), dtypes.dtype_code(t.dtype), _auto_name("lbcast", name, self._version))
        return t

    def gather_cpu(self, t: torch.Tensor, name: str = "") -> Optional[torch.Tensor]:
        out = t.new_empty((self.size,) + tuple(t.shape)) if self.rank == 0 else t.new_empty((0,))
        self.peer.gather(t.data_ptr(), out.data_ptr(), t.numel(), dtypes.dtype_code(t.dtype), _auto_name("gather", name, self._version))
        return out if self.rank == 0 else None

    def all_gather_cpu(self, t: torch.Tensor, name: str = "") -> torch.Tensor:
        out = t.new_empty((self.size,) + tuple(t.shape))
        self.peer.all_gather(t.data_ptr(), out.data_ptr(), t.numel(), dtypes.dtype_code(t.dtype), _auto_name("allgather", name, self._version))
        return out

    def all_reduce_int_max(self, x: int) -> int:
        return int(self.peer.all_reduce_int_max(int(x)))

    @staticmethod
    def wait(handle: int) -> None:
        _kfrt.wait(handle)

    @staticmethod
    def wait_all(handles: List[int]) -> None:
        _kfrt.wait_all(list(handles))

    # ---- adaptation / monitoring ----------------------------------------------------------------------
    def set_tree(self, father: List[int]) -> None:
        self.peer.set_tree([int(x) for x in father])

    def peer_latencies(self) -> List[float]:
        return list(self.peer.peer_latencies())

    def egress_rates(self) -> List[float]:
        return list(self.peer.egress_rates())

    def calc_stats(self) -> None:
        self.peer.calc_stats()

    def log_stats(self) -> None:
        self.peer.log_stats()

    def print_strategy_stats(self) -> None:
        self.peer.print_strategy_stats()

    def check_interference(self) -> bool:
        return bool(self.peer.check_interference())

    # ---- elastic --------------------------------------------------------------------------------------
    def resize(self, n: Optional[int] = None) -> Tuple[bool, bool]:
        changed, detached = self.peer.resize_cluster(int(n)) if n is not None else self.peer.resize_cluster_from_url()
        if changed and not detached:
            self._refresh()
        return bool(changed), bool(detached)

    def propose_new_size(self, n: int) -> None:
        self.peer.propose_new_size(int(n))

    def recover(self, timeout_s: float = 30.0) -> Tuple[bool, bool]:
        """Adopt the cluster the runners announced after a worker died (see `Peer::recover`). Returns (changed, detached)."""
        changed, detached = self.peer.recover(float(timeout_s))
        if changed and not detached:
            self._refresh()
        return bool(changed), bool(detached)

    def failure_pending(self) -> bool:
        return bool(self.peer.failure_pending())

    # ---- p2p store ------------------------------------------------------------------------------------
    def save(self, name: str, t: torch.Tensor, version: Optional[str] = None) -> None:
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        if version is None:
            self.peer.save(name, t.data_ptr(), nbytes)
        else:
            self.peer.save_version(str(version), name, t.data_ptr(), nbytes)

    def request(self, rank: int, name: str, out: torch.Tensor, version: Optional[str] = None) -> bool:
        assert out.is_contiguous() and not out.is_cuda
        return bool(self.peer.request(int(rank), "" if version is None else str(version), name, out.data_ptr(),
                                      out.numel() * out.element_size()))

    def close(self) -> None:
        try:
            if hasattr(self.peer, "set_lost_callback"):
                self.peer.set_lost_callback(None)
            self.peer.close()
        except Exception:
            pass


def config_server_url() -> str:
    return os.environ.get("KUNGFU_CONFIG_SERVER", "")
