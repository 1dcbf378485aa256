"""Device plane: symmetric, peer-mapped HBM arenas on every rank.

`SymmetricArena` wraps the native `PeerMesh` (csrc/mesh/peer_mesh.cpp). Every rank creates arenas in the
same order with the same sizes, so a byte offset names the same logical buffer on all ranks — that
offset is all a kernel needs to address its peers (`MeshArgs.data[r] + off`).

Replaces the reference's `NCCLHelper` / `NCCLController` pair (srcs/cpp/include/kungfu/nccl/helper.hpp:14-27,
controller.hpp:9-33): there is no communicator object, only mapped memory + signal pads.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch

from ..utils import dtypes


def _load_ext():
    try:
        from .. import _kfcuda  # type: ignore[attr-defined]

        return _kfcuda
    except ImportError as e:  # fail loudly: a silent PyTorch fallback would hide a missing native build
        raise ImportError(
            "kungfu_b200: the CUDA extension `_kfcuda` is not built. Run `python -m kungfu_b200._build` "
            "(or `__graft_entry__.build()`) first.") from e


class _CudaView:
    """Zero-copy torch view of raw device memory via the CUDA array interface."""

    def __init__(self, ptr: int, nbytes: int, owner) -> None:
        self._owner = owner  # keeps the arena alive while tensors view it
        self.__cuda_array_interface__ = {
            "shape": (nbytes,),
            "typestr": "|u1",
            "data": (ptr, False),
            "version": 3,
            "strides": None,
        }


class SymmetricArena:
    """One symmetric allocation + signal pads + (optional) NVLS multicast alias."""

    def __init__(self, plane: "GpuPlane", nbytes: int, name: str = "arena", want_multicast: bool = True) -> None:
        ext = _load_ext()
        self.plane = plane
        self.name = name
        self.ext = ext
        rt = plane.runtime
        # an elastic job may grow beyond one rank later: its arenas must be exportable (VMM) from the start
        elastic = bool(os.environ.get("KUNGFU_CONFIG_SERVER")) or "KUNGFU_SELF_SPEC" in os.environ
        mode = os.environ.get("KUNGFU_MESH_MODE", "auto" if (rt.size > 1 or elastic) else "local")
        self.mesh = ext.PeerMesh(plane.device_index, int(nbytes), mode)
        self.mesh.set_timeout_ms(float(os.environ.get("KUNGFU_DEVICE_TIMEOUT_MS", "20000")))
        self._cursor = 0
        self._regions: Dict[str, Tuple[int, int]] = {}
        self._zero_on_connect: List[Tuple[int, int]] = []
        self._want_multicast = want_multicast
        self._epoch = None  # membership this arena is connected for; connection is lazy (see `ensure`)
        base = _CudaView(self.mesh.local_ptr, self.mesh.arena_bytes, self)
        with torch.cuda.device(plane.device_index):
            self._bytes = torch.as_tensor(base, device=torch.device("cuda", plane.device_index))

    def _connect(self, want_multicast: bool) -> None:
        plane = self.plane
        want_mc = want_multicast and os.environ.get("KUNGFU_DISABLE_NVLS", "0") != "1"
        # peer memory exists only inside one NVLink domain: the mesh spans the ranks of THIS host (all of them in a
        # single-host job); ranks on other hosts are reached through the hierarchical paths of ops/collective.py
        self.mesh.connect(plane.group_rank, plane.group_size, plane.group_all_gather_bytes, want_mc)

    def ensure(self) -> "SymmetricArena":
        """Connect (or re-connect after a membership change) to the current peers. COLLECTIVE when it has work to do:
        every rank reaches the first use of an arena in the same program order, and after a resize all members call
        `GpuPlane.reconnect_all()` from `sync_after_resize`. Constructors never communicate, so late joiners can build
        their objects without the survivors taking part. Local memory and every tensor viewing it survive."""
        key = self.plane.membership_key()
        if self._epoch != key:
            torch.cuda.synchronize(self.plane.device_index)
            for off, size in self._zero_on_connect:
                # regions validated by operation numbers that restart with the membership (LL slots): stale flags must be gone
                # before the connect's closing barrier lets any peer push again
                self._bytes[off:off + size].zero_()
            torch.cuda.synchronize(self.plane.device_index)
            self._connect(self._want_multicast)
            self._epoch = key
        return self

    # ---- allocation -------------------------------------------------------------------------------
    @property
    def nbytes(self) -> int:
        return self.mesh.arena_bytes

    @property
    def has_multicast(self) -> bool:
        return bool(self.mesh.has_multicast)

    def alloc(self, nbytes: int, align: int = 256, name: Optional[str] = None) -> int:
        """Bump-allocate `nbytes`; returns the byte offset (identical on every rank)."""
        if name is not None and name in self._regions:
            off, size = self._regions[name]
            if size < nbytes:
                raise ValueError(f"region {name!r} exists with {size} < {nbytes} bytes")
            return off
        off = (self._cursor + align - 1) // align * align
        end = off + (nbytes + 15) // 16 * 16
        if end > self.nbytes:
            raise MemoryError(f"kungfu_b200: arena {self.name!r} exhausted ({end} > {self.nbytes} bytes)")
        self._cursor = end
        if name is not None:
            self._regions[name] = (off, end - off)
        return off

    def view(self, off: int, numel: int, dtype: torch.dtype) -> torch.Tensor:
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        return self._bytes[off:off + nbytes].view(dtype)

    def bytes_view(self, off: int, nbytes: int) -> torch.Tensor:
        return self._bytes[off:off + nbytes]

    def check(self) -> None:
        """Raise if a device-side wait timed out (dead / diverged peer)."""
        st = self.mesh.status()
        if st[0] != 0:
            info = (f"code={st[0]} rank={st[1] >> 16} waiting_for_peer={st[1] & 0xffff} timeouts={st[2]} "
                    f"want_seq={st[3]} seen_seq={st[4]}")
            self.mesh.clear_status()
            if st[0] == 3:
                raise RuntimeError("kungfu_b200: device-side waits aborted by the host, a peer lost: " + info)
            raise RuntimeError("kungfu_b200: device-side barrier timed out: " + info)

    def close(self) -> None:
        self._bytes = None  # type: ignore[assignment]
        try:
            self.mesh.disconnect()
        except Exception:
            pass


class GpuPlane:
    """Per-process GPU context: device selection, comm streams, the default staging arena."""

    def __init__(self, runtime) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("kungfu_b200: no CUDA device available")
        self.runtime = runtime
        self.ext = _load_ext()
        self.device_index = runtime.cuda_index()
        if self.device_index < 0:
            raise RuntimeError("kungfu_b200: this worker has no GPU assigned")
        torch.cuda.set_device(self.device_index)
        self.device = torch.device("cuda", self.device_index)
        # high-priority streams so that comm CTAs are scheduled as soon as SMs drain
        self.comm_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.side_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self._arenas: List[SymmetricArena] = []
        self.engines: list = []  # FlatEngines living on this plane (told about membership changes)
        self._staging: Optional[SymmetricArena] = None
        self._staging_off = 0
        self._staging_bytes = 0
        self.sm_count = torch.cuda.get_device_properties(self.device_index).multi_processor_count
        # shrink-and-continue: the moment a runner announces a dead worker, every device-side wait is told to give up (the
        # host-mapped status word) instead of sitting out KUNGFU_DEVICE_TIMEOUT_MS; `python.recover()` clears it again
        peer = getattr(runtime.control, "peer", None)
        if peer is not None and hasattr(peer, "set_lost_callback"):
            peer.set_lost_callback(self.abort_device_waits)

    def abort_device_waits(self) -> None:
        for a in list(self._arenas):
            try:
                a.mesh.abort()
            except Exception:
                pass

    def membership_key(self):
        rt = self.runtime
        ver = getattr(rt.control, "cluster_version", None)
        return (rt.size, rt.rank, ver() if ver else 0)

    # ---- the NVLink group of this rank ---------------------------------------------------------------------------------
    # A job on one host: group == world. A job spanning hosts (reference: `kungfu-run -H a:8,b:8`, CrossAllReduceGpu in
    # srcs/cpp/src/nccl/controller.cpp): the group is this host's ranks (contiguous in the peer list, which is host-major),
    # kernels run inside the group and the host dimension goes through the control plane's cross-host collectives.
    @property
    def spans_hosts(self) -> bool:
        rt = self.runtime
        forced = os.environ.get("KUNGFU_FORCE_GROUP_SIZE")  # testing: split one box into several "hosts"
        ls = int(forced) if forced else int(rt.local_size)
        return 1 <= ls < rt.size

    @property
    def group_size(self) -> int:
        rt = self.runtime
        if not self.spans_hosts:
            return rt.size
        forced = os.environ.get("KUNGFU_FORCE_GROUP_SIZE")
        return int(forced) if forced else int(rt.local_size)

    @property
    def group_rank(self) -> int:
        rt = self.runtime
        if not self.spans_hosts:
            return rt.rank
        forced = os.environ.get("KUNGFU_FORCE_GROUP_SIZE")
        return rt.rank % int(forced) if forced else int(rt.local_rank)

    @property
    def group_base(self) -> int:
        return self.runtime.rank - self.group_rank

    def group_all_gather_bytes(self, b) -> List[bytes]:
        """Out-of-band all-gather restricted to the NVLink group (built on the world all-gather: every rank of every
        host reaches the same connect step in the same program order)."""
        rt = self.runtime
        everyone = rt.control.all_gather_bytes(b)
        if not self.spans_hosts:
            return everyone
        base, n = self.group_base, self.group_size
        if base < 0 or base + n > len(everyone):
            raise RuntimeError(f"kungfu_b200: ranks of one host are not contiguous (rank {rt.rank}, local {self.group_rank}/{n})")
        return everyone[base:base + n]

    def new_arena(self, nbytes: int, name: str, want_multicast: bool = True) -> SymmetricArena:
        names = {a.name for a in self._arenas}
        base, k = name, 1
        while name in names:  # unique, deterministic names: the reconnect order after a resize is by name
            k += 1
            name = f"{base}#{k}"
        a = SymmetricArena(self, nbytes, name, want_multicast)
        self._arenas.append(a)
        return a

    def staging(self) -> Tuple[SymmetricArena, int, int]:
        """Default arena used to stage arbitrary tensors for the generic collectives."""
        if self._staging is None:
            mb = int(os.environ.get("KUNGFU_STAGING_MB", "64"))
            self.ll_max = int(os.environ.get("KUNGFU_LL_MAX", str(64 << 10)))  # payload bytes of the barrier-free LL one-shot
            ll_bytes = self.ext.allreduce_ll_slot_bytes(self.ll_max)
            self._staging = self.new_arena(((mb + 2) << 20) + ll_bytes + (1 << 16), "staging")
            self._staging_bytes = mb << 20
            self._staging_off = self._staging.alloc(self._staging_bytes, name="stage")
            self.ll_off = self._staging.alloc(ll_bytes, align=4096, name="ll")
            self._staging._zero_on_connect.append((self.ll_off, ll_bytes))
        self._staging.ensure()
        return self._staging, self._staging_off, self._staging_bytes

    def reconnect_all(self) -> None:
        """After a membership change: all members (survivors and newcomers) re-discover their peers and rebuild the
        mappings of every arena they have in common, in name order (K4 / `PeerMesh::connect`)."""
        rt = self.runtime
        self.staging()
        mine = sorted(a.name for a in self._arenas)
        if rt.size > 1:
            lists = [set(b.decode().split("\n")) for b in rt.control.all_gather_bytes("\n".join(mine).encode())]
            common = set.intersection(*lists)
        else:
            common = set(mine)
        for a in sorted(self._arenas, key=lambda a: a.name):
            if a.name in common:
                a.ensure()

    def topology(self) -> dict:
        st, _, _ = self.staging()
        return st.mesh.topology()

    def close(self) -> None:
        try:
            torch.cuda.synchronize(self.device_index)
        except Exception:
            pass
        for a in self._arenas:
            a.close()
        self._arenas.clear()
        self._staging = None


_DT = dtypes
