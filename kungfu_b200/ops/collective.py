"""Collective operations: GPU tensors go through our sm_100a kernels over peer-mapped HBM, CPU tensors
through the native CPU runtime (graph strategies over sockets).

API parity: `kungfu.torch.ops.{all_reduce_fn,inplace_all_reduce_op,inplace_all_reduce_async_op,wait_handle,
wait_all_handles,all_gather,broadcast_parameters}` (srcs/python/kungfu/torch/ops/collective.py:8-55) and the TF
ops `barrier/consensus/broadcast/all_reduce/all_gather` (srcs/python/kungfu/tensorflow/ops/collective.py:8-133).

Algorithm table (KUNGFU_CONFIG_ALGO = auto | ll | oneshot | twoshot | nvls):
  <= 64 KiB (8-byte aligned) : LL one-shot — flag-carrying pushes straight from / into the user tensor, no barrier, no staging
  <= oneshot_max bytes  : one-shot (every rank reduces the whole message; latency path)
  larger                : NVLS multimem reduce+multicast when the mesh has a multicast mapping and the
                          dtype/op is switch-reducible, else two-shot P2P.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional

import torch

from ..runtime import get_runtime
from ..utils import dtypes


@dataclass
class AlgoConfig:
    algo: str = os.environ.get("KUNGFU_CONFIG_ALGO", "auto").lower()
    oneshot_max: int = int(os.environ.get("KUNGFU_ONESHOT_MAX", str(256 << 10)))
    threads: int = int(os.environ.get("KUNGFU_COMM_THREADS", "512"))
    max_blocks: int = int(os.environ.get("KUNGFU_COMM_BLOCKS", "32"))
    vec_per_thread: int = 2


CONFIG = AlgoConfig()

_ALGO_CODE = {"auto": 0, "oneshot": 1, "twoshot": 2, "nvls": 3}


def pick_algo(arena, nbytes: int, dtype_code: int, op_code: int, inplace: bool, force: Optional[str] = None) -> int:
    """Returns the kernel algorithm code for a message."""
    ext = arena.ext
    name = (force or CONFIG.algo).lower()
    if name in ("nccl", "ll"):
        name = "auto"
    if name != "auto":
        code = _ALGO_CODE[name]
        if code == ext.ALGO_NVLS and not (arena.has_multicast and ext.nvls_supported(dtype_code, op_code)):
            code = ext.ALGO_TWOSHOT
        if code == ext.ALGO_ONESHOT and inplace:
            code = ext.ALGO_TWOSHOT
        return code
    if not inplace and nbytes <= CONFIG.oneshot_max:
        return ext.ALGO_ONESHOT
    if arena.has_multicast and ext.nvls_supported(dtype_code, op_code):
        return ext.ALGO_NVLS
    return ext.ALGO_TWOSHOT


def pick_blocks(nbytes: int, nranks: int, sliced: bool, max_blocks: Optional[int] = None) -> int:
    """Grid size: enough CTAs to keep NVLink busy, few enough to leave the SMs to backward."""
    mb = max_blocks or CONFIG.max_blocks
    nvec = (nbytes + 15) // 16
    work = (nvec + nranks - 1) // nranks if sliced else nvec
    per_block = CONFIG.threads * CONFIG.vec_per_thread
    return max(1, min(mb, (work + per_block - 1) // per_block))


class Handle:
    """Completion handle of an asynchronous collective (reference: `HandleManager`,
    srcs/cpp/include/kungfu/utils/handler_manager.hpp:40-83 — busy-waiting there, a CUDA event here)."""

    __slots__ = ("event", "post", "arena")

    def __init__(self, event=None, post=None, arena=None):
        self.event = event
        self.post = post
        self.arena = arena

    def wait(self) -> None:
        if self.event is not None:
            torch.cuda.current_stream().wait_event(self.event)
            self.event = None
        if self.post is not None:
            self.post()
            self.post = None


def _run_on_comm(plane, fn, tensors=(), wait: bool = True):
    """Run `fn(stream)` on the plane's comm stream, ordered after the caller's stream. All staged
    collectives share the staging slot and signal channel 0, so they are serialised on ONE stream."""
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(cur)
    plane.comm_stream.wait_event(ev)
    with torch.cuda.stream(plane.comm_stream):
        fn(plane.comm_stream)
    for t in tensors:
        t.record_stream(plane.comm_stream)
    done = torch.cuda.Event()
    done.record(plane.comm_stream)
    if wait:
        cur.wait_event(done)
    return done


def _gpu_all_reduce_(x: torch.Tensor, op_code: int, average: bool, stream: Optional[torch.cuda.Stream] = None,
                     algo: Optional[str] = None) -> None:
    """In-place all-reduce of an arbitrary contiguous CUDA tensor, staged through the symmetric arena.
    Must be called with `stream` = the plane's comm stream (see `_run_on_comm`)."""
    rt = get_runtime()
    plane = rt.gpu
    arena, off, cap = plane.staging()
    ext = plane.ext
    dt = dtypes.dtype_code(x.dtype)
    flat = x.view(-1)
    esize = x.element_size()
    st = stream or torch.cuda.current_stream()
    scale = 1.0 / rt.size if average else 1.0
    flags = ext.FLAG_SCALE if average else 0
    if average and not x.dtype.is_floating_point:
        raise TypeError("kungfu_b200: averaging all-reduce needs a floating point tensor")
    chunk_elems = (cap // esize) // 8 * 8
    n = flat.numel()
    _count_gpu_bytes("all_reduce", n * esize)
    if plane.spans_hosts:
        _gpu_all_reduce_hier_(plane, flat, op_code, average, st, algo)
        return
    nbytes = n * esize
    # every rank pushes its whole (flag-doubled) message to N-1 peers: beyond ~128 KiB of pushes per rank the copy + barrier one-shot
    # is faster (measured, profiles/allreduce_bw_8gpu.json: 64 KiB at N=8 LL 12.8 us vs one-shot 9.7 us; at N=2 LL wins up to 64 KiB)
    forced_ll = (algo or CONFIG.algo).lower() == "ll"
    ll_cap = plane.ll_max if forced_ll else min(plane.ll_max, (128 << 10) // max(1, plane.group_size - 1))
    if (nbytes <= ll_cap and nbytes % 8 == 0 and (algo or CONFIG.algo).lower() in ("auto", "ll")
            and os.environ.get("KUNGFU_DISABLE_LL", "0") != "1"):
        # latency path: flag-carrying pushes straight from / into the user tensor — no staging copy, no barrier. The choice
        # depends on the message size only (identical on every rank); a misaligned view goes through an aligned temporary.
        with torch.cuda.stream(st):
            buf = flat if flat.data_ptr() % 8 == 0 else flat.clone()
            ext.allreduce_ll(arena.mesh, 0, buf.data_ptr(), buf.data_ptr(), nbytes, plane.ll_off, plane.ll_max, dt, op_code, scale, flags,
                             st.cuda_stream)
            if buf is not flat:
                flat.copy_(buf)
        return
    with torch.cuda.stream(st):
        for lo in range(0, n, chunk_elems):
            hi = min(n, lo + chunk_elems)
            nb = (hi - lo) * esize
            nb_pad = (nb + 15) // 16 * 16
            stage = arena.bytes_view(off, nb_pad)
            stage[:nb].copy_(flat[lo:hi].view(torch.uint8), non_blocking=True)
            if nb_pad != nb:
                stage[nb:].zero_()
            a = pick_algo(arena, nb_pad, dt, op_code, inplace=False, force=algo)
            if a == ext.ALGO_ONESHOT and nb_pad == nb and flat[lo:hi].data_ptr() % 16 == 0:
                # results go straight into the user tensor
                blocks = pick_blocks(nb_pad, rt.size, sliced=False)
                ext.allreduce(arena.mesh, 0, a, off, flat[lo:hi].data_ptr(), nb_pad, dt, op_code, scale, 0, flags, blocks,
                              CONFIG.threads, st.cuda_stream)
            else:
                if a == ext.ALGO_ONESHOT:
                    a = ext.ALGO_TWOSHOT
                blocks = pick_blocks(nb_pad, rt.size, sliced=True)
                ext.allreduce(arena.mesh, 0, a, off, 0, nb_pad, dt, op_code, scale, 0, flags, blocks, CONFIG.threads,
                              st.cuda_stream)
                flat[lo:hi].view(torch.uint8).copy_(stage[:nb], non_blocking=True)


_OPNAME = {0: "sum", 1: "min", 2: "max", 3: "prod"}
_metric_add = None


def _count_gpu_bytes(op: str, nbytes: int) -> None:
    """`gpu_collective_bytes_total{op=...}` on this worker's /metrics: payload bytes handed to the staged device collectives
    (the fused optimizer engines run inside CUDA graphs and account for themselves per step, not per launch)."""
    global _metric_add
    if _metric_add is None:
        try:
            from .. import _kfrt  # type: ignore[attr-defined]

            _metric_add = _kfrt.metric_add
        except ImportError:
            _metric_add = lambda *_: None  # noqa: E731
    _metric_add('gpu_collective_bytes_total{op="%s"}' % op, float(nbytes))


def _group_masters(plane) -> List[int]:
    """World ranks of the first rank of every NVLink group (COLLECTIVE on first use per membership)."""
    key = plane.membership_key()
    if getattr(plane, "_masters_key", None) != key:
        rt = plane.runtime
        bases = [int(b) for b in rt.control.all_gather_bytes(str(plane.group_base).encode())]
        plane._masters = sorted(set(bases))
        plane._masters_key = key
    return plane._masters


def _pinned(plane, nbytes: int) -> torch.Tensor:
    buf = getattr(plane, "_pinned_buf", None)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        plane._pinned_buf = buf
    return buf


def _cross_hosts_(plane, host: torch.Tensor, op_code: int, tag: str = "") -> None:
    """All-reduce a host tensor among the group masters (one rank per host) over the control plane. Reference:
    `CrossAllReduceGpu` stages through host memory the same way (srcs/cpp/src/nccl/controller.cpp, K5 in SURVEY.md)."""
    rt = plane.runtime
    masters = _group_masters(plane)
    if len(masters) < 2:
        return
    ctl = rt.control
    if "KUNGFU_FORCE_GROUP_SIZE" not in os.environ and hasattr(ctl, "cross_all_reduce_cpu_"):
        # real hosts: the session's cross-host strategies (ring / binary tree over the hosts' first ranks, chosen by -strategy)
        ctl.cross_all_reduce_cpu_(host, _OPNAME[op_code], tag)
    elif hasattr(ctl, "subset_all_reduce_cpu_"):
        forest = list(range(rt.size))
        for m in masters:
            forest[m] = masters[0]
        ctl.subset_all_reduce_cpu_(host, forest, _OPNAME[op_code], tag)
    elif op_code == dtypes.SUM:
        ctl.all_reduce_cpu_(host, "sum", tag)  # control planes without subset support: everyone joins, non-masters add zeros
    else:
        raise NotImplementedError("kungfu_b200: this control plane has no subset all-reduce for min/max/prod across hosts")


def _gpu_all_reduce_hier_(plane, flat: torch.Tensor, op_code: int, average: bool, st, algo: Optional[str]) -> None:
    """All-reduce for jobs that span hosts: (1) our kernel inside the NVLink group, (2) the group's first rank takes the
    partial result across hosts through pinned memory and the control plane, (3) broadcast kernel inside the group."""
    rt = plane.runtime
    arena, off, cap = plane.staging()
    ext = plane.ext
    dt = dtypes.dtype_code(flat.dtype)
    esize = flat.element_size()
    gsize, grank = plane.group_size, plane.group_rank
    masters = _group_masters(plane)
    subset = hasattr(rt.control, "subset_all_reduce_cpu_")
    ver = getattr(rt.control, "cluster_version", None)
    key_version = ver() if ver else 0
    if getattr(plane, "_hier_version", None) != key_version:  # numbering restarts with the membership, like every op name
        plane._hier_version, plane._hier_seq = key_version, 0
    chunk_elems = (cap // esize) // 8 * 8
    n = flat.numel()
    with torch.cuda.stream(st):
        for lo in range(0, n, chunk_elems):
            hi = min(n, lo + chunk_elems)
            nb = (hi - lo) * esize
            nb_pad = (nb + 15) // 16 * 16
            # every rank numbers the hierarchical chunks identically; the host-side ops below are named after that number, so that
            # ops only the hosts' first ranks perform cannot shift anybody's automatic op names
            plane._hier_seq = getattr(plane, "_hier_seq", 0) + 1
            tag = f"kungfu::hier::v{key_version}::{plane._hier_seq}"
            stage = arena.bytes_view(off, nb_pad)
            stage[:nb].copy_(flat[lo:hi].view(torch.uint8), non_blocking=True)
            if nb_pad != nb:
                stage[nb:].zero_()
            if gsize > 1:
                a = pick_algo(arena, nb_pad, dt, op_code, inplace=True, force=algo)
                if a == ext.ALGO_ONESHOT:
                    a = ext.ALGO_TWOSHOT
                ext.allreduce(arena.mesh, 0, a, off, 0, nb_pad, dt, op_code, 1.0, 0, 0, pick_blocks(nb_pad, gsize, sliced=True),
                              CONFIG.threads, st.cuda_stream)
            if len(masters) > 1:
                joins = grank == 0 or not subset
                if joins:
                    host = _pinned(plane, nb_pad)[:nb]
                    if grank == 0:
                        host.copy_(stage[:nb], non_blocking=True)
                        st.synchronize()
                    else:
                        host.zero_()
                    _cross_hosts_(plane, host.view(flat.dtype), op_code, tag + "::x")
                    if grank == 0:
                        stage[:nb].copy_(host, non_blocking=True)
                if gsize > 1:
                    # the group waits for its master on the host (not spinning on the device while the master is on the network)
                    sync = getattr(rt.control, "local_broadcast_cpu_", None)
                    if sync is not None:
                        sync(torch.zeros(1, dtype=torch.int32), tag + "::s")
                    algo_b = ext.ALGO_NVLS if (arena.has_multicast and CONFIG.algo in ("auto", "nvls")) else ext.ALGO_TWOSHOT
                    ext.broadcast(arena.mesh, 0, algo_b, off, nb_pad, 0, pick_blocks(nb_pad, 1, sliced=False), CONFIG.threads,
                                  st.cuda_stream)
            out = flat[lo:hi]
            out.view(torch.uint8).copy_(stage[:nb], non_blocking=True)
            if average:
                out.mul_(1.0 / rt.size)


def all_reduce_(x: torch.Tensor, op=None, name: str = "") -> torch.Tensor:
    """In-place all-reduce. `op`: sum (default) | min | max | prod | avg."""
    code, avg = dtypes.op_code(op)
    rt = get_runtime()
    if rt.size == 1:
        return x
    if x.is_cuda:
        plane = rt.gpu
        if not x.is_contiguous():
            y = x.contiguous()
            _run_on_comm(plane, lambda s: _gpu_all_reduce_(y, code, avg, stream=s), (y,))
            x.copy_(y)
        else:
            _run_on_comm(plane, lambda s: _gpu_all_reduce_(x, code, avg, stream=s), (x,))
        return x
    opname = {0: "sum", 1: "min", 2: "max", 3: "prod"}[code]
    if x.is_contiguous():
        rt.control.all_reduce_cpu_(x, opname, name)
        if avg:
            x.div_(rt.size)
    else:
        y = x.contiguous()
        rt.control.all_reduce_cpu_(y, opname, name)
        if avg:
            y.div_(rt.size)
        x.copy_(y)
    return x


def all_reduce(x: torch.Tensor, op=None, name: str = "") -> torch.Tensor:
    """Out-of-place all-reduce (`op`: sum | min | max | prod | avg). Reference: ops/collective.py `all_reduce`, torch `all_reduce_fn`."""
    return all_reduce_(x.clone(), op, name)


def subset_all_reduce(x: torch.Tensor, topology, op=None, name: str = "") -> torch.Tensor:
    """All-reduce inside every tree of the forest `topology` (father array: `topology[i] == i` marks a root; a root
    without children keeps its own value). Reference: `subset_all_reduce(t, topology)` (srcs/python/kungfu/tensorflow/ops/
    collective.py, kernel `KungfuSubsetAllReduce`). Members of multi-node trees must all call it; others may skip it."""
    code, avg = dtypes.op_code(op)
    rt = get_runtime()
    forest = [int(v) for v in (topology.tolist() if isinstance(topology, torch.Tensor) else topology)]
    if len(forest) != rt.size:
        raise ValueError(f"subset_all_reduce: topology has {len(forest)} entries for {rt.size} peers")
    if rt.size == 1:
        return x.clone()
    f = getattr(rt.control, "subset_all_reduce_cpu_", None)
    if f is None:
        raise NotImplementedError("kungfu_b200: subset_all_reduce needs the native control plane")
    host = x.detach().to("cpu").contiguous().clone()
    f(host, forest, _OPNAME[code], name)
    if avg:
        root = rt.rank
        while forest[root] != root:
            root = forest[root]
        members = 0
        for i in range(rt.size):
            j = i
            while forest[j] != j:
                j = forest[j]
            members += j == root
        host = host / members if host.dtype.is_floating_point else host // members
    return host.to(x.device)


def all_reduce_async_(x: torch.Tensor, name: str = "", op=None) -> Handle:
    """Asynchronous in-place all-reduce on the comm stream; `Handle.wait()` makes the current stream
    wait for it (no host blocking)."""
    code, avg = dtypes.op_code(op)
    rt = get_runtime()
    if rt.size == 1:
        return Handle()
    if not x.is_cuda:
        # Host tensors: the native collective releases the GIL, so a small FIFO executor keeps several named collectives in
        # flight (the reference's TF ops run concurrently the same way; per-tensor latency, not bandwidth, dominates a model
        # with 160 small tensors). The op name is fixed HERE, in program order, so that every rank pairs the same tensors.
        if getattr(rt.control, "kind", "") != "native" or os.environ.get("KUNGFU_CPU_ASYNC", "1") == "0":
            all_reduce_(x, op, name)
            return Handle()
        fut = _cpu_executor().submit(all_reduce_, x, op, name or _async_name(rt))
        return Handle(post=fut.result)
    plane = rt.gpu
    done = _run_on_comm(plane, lambda s: _gpu_all_reduce_(x, code, avg, stream=s), (x,), wait=False)
    return Handle(event=done)


_cpu_pool = None
_async_seq = {"version": None, "n": 0}


def _cpu_executor():
    global _cpu_pool
    if _cpu_pool is None:
        from concurrent.futures import ThreadPoolExecutor

        _cpu_pool = ThreadPoolExecutor(max_workers=int(os.environ.get("KUNGFU_CPU_ASYNC_THREADS", "8")), thread_name_prefix="kf-cpu-async")
    return _cpu_pool


def _async_name(rt) -> str:
    """Deterministic name for an unnamed asynchronous op; numbering restarts with the cluster version (see peer._auto_name)."""
    ver = getattr(rt.control, "cluster_version", None)
    v = ver() if ver else 0
    if _async_seq["version"] != v:
        _async_seq["version"], _async_seq["n"] = v, 0
    _async_seq["n"] += 1
    return f"async-allreduce::v{v}::{_async_seq['n']}"


def all_reduce_group_async_(ts: List[torch.Tensor], op=None) -> Handle:
    """All-reduce a LIST of CUDA tensors in place with ONE hop onto the comm stream and one completion event: the per-tensor
    cost is one kernel launch (small tensors: the barrier-free LL kernel reading / writing the user tensor directly). The
    reference's `group_all_reduce` issues one named op per tensor (ops/collective.py:67-69); their list order is the order on
    every rank here too. Host tensors fall back to per-tensor ops."""
    code, avg = dtypes.op_code(op)
    rt = get_runtime()
    ts = [t for t in ts if t is not None]
    if rt.size == 1 or not ts:
        return Handle()
    if not all(t.is_cuda and t.is_contiguous() for t in ts):
        hs = [all_reduce_async_(t, f"group::{i}", op) for i, t in enumerate(ts)]
        return Handle(post=lambda: wait_all_handles(hs))
    plane = rt.gpu

    def run(st):
        for t in ts:
            _gpu_all_reduce_(t, code, avg, stream=st)

    return Handle(event=_run_on_comm(plane, run, ts, wait=False))


def wait_handle(h: Handle) -> None:
    """Make the current stream (CUDA) or thread (host tensors) wait for one asynchronous collective."""
    h.wait()


def wait_all_handles(hs: List[Handle]) -> None:
    """`wait_handle` for every handle of a list (reference: torch/ops/collective.py wait_all_handles)."""
    for h in hs:
        h.wait()


def broadcast_(x: torch.Tensor, root: int = 0, name: str = "") -> torch.Tensor:
    """In-place broadcast of `x` from rank `root` (NVLS multicast push or P2P pull on GPUs, graph broadcast on the host plane)."""
    rt = get_runtime()
    if rt.size == 1:
        return x
    if x.is_cuda and rt.gpu.spans_hosts:
        h = x.detach().to("cpu")
        broadcast_(h, root, name)
        x.copy_(h)
        return x
    if x.is_cuda:
        plane = rt.gpu
        arena, off, cap = plane.staging()
        ext = plane.ext
        flat = x.contiguous().view(-1).view(torch.uint8)
        n = flat.numel()

        def run(st):
            for lo in range(0, n, cap):
                hi = min(n, lo + cap)
                nb = hi - lo
                nb_pad = (nb + 15) // 16 * 16
                stage = arena.bytes_view(off, nb_pad)
                if rt.rank == root:
                    stage[:nb].copy_(flat[lo:hi], non_blocking=True)
                algo = ext.ALGO_NVLS if (arena.has_multicast and CONFIG.algo in ("auto", "nvls")) else ext.ALGO_TWOSHOT
                blocks = pick_blocks(nb_pad, 1, sliced=False)
                ext.broadcast(arena.mesh, 0, algo, off, nb_pad, root, blocks, CONFIG.threads, st.cuda_stream)
                if rt.rank != root:
                    flat[lo:hi].copy_(stage[:nb], non_blocking=True)

        _run_on_comm(plane, run, (flat,))
        if not x.is_contiguous():
            x.copy_(flat.view(x.dtype).view(x.shape))
        return x
    native = getattr(rt.control, "broadcast_cpu_", None)
    if native is not None and root == 0 and x.is_contiguous():
        native(x, name)  # graph broadcast from rank 0, zero-copy on the tensor's memory
        return x
    data = x.contiguous().view(-1).view(torch.uint8).numpy().tobytes() if rt.rank == root else b""
    got = rt.control.broadcast_bytes(data, root)
    if rt.rank != root:
        x.copy_(torch.frombuffer(bytearray(got), dtype=x.dtype).view(x.shape))
    return x


def broadcast_async_(x: torch.Tensor, name: str = "", root: int = 0) -> Handle:
    """Asynchronous in-place broadcast (reference: torch/ops/collective.py `inplace_broadcast_async_op`, used by
    `broadcast_parameters` to keep every tensor of a state dict in flight). Host tensors go through the same FIFO executor as the
    asynchronous all-reduce (named in program order); device tensors are already stream-ordered — `broadcast_` enqueues on the comm
    stream and makes the current stream wait, the host never blocks — so their handle is complete on return."""
    rt = get_runtime()
    if rt.size == 1:
        return Handle()
    if x.is_cuda or getattr(rt.control, "kind", "") != "native" or os.environ.get("KUNGFU_CPU_ASYNC", "1") == "0":
        broadcast_(x, root, name)
        return Handle()
    fut = _cpu_executor().submit(broadcast_, x, root, name or _async_name(rt).replace("allreduce", "broadcast"))
    return Handle(post=fut.result)


def all_gather(x: torch.Tensor, name: str = "") -> torch.Tensor:
    """Returns a tensor of shape [np, *x.shape] (reference: torch/ops/collective.py:50-55)."""
    rt = get_runtime()
    n = rt.size
    out = x.new_empty((n,) + tuple(x.shape))
    if n == 1:
        out[0].copy_(x)
        return out
    if x.is_cuda and rt.gpu.spans_hosts:
        out.copy_(all_gather(x.detach().to("cpu"), name))
        return out
    if x.is_cuda:
        plane = rt.gpu
        arena, off, cap = plane.staging()
        ext = plane.ext
        flat = x.contiguous().view(-1).view(torch.uint8)
        nb = flat.numel()
        nb_pad = (nb + 15) // 16 * 16
        if nb_pad * (n + 1) > cap:
            raise ValueError("kungfu_b200: all_gather message exceeds the staging arena (raise KUNGFU_STAGING_MB)")
        in_off, out_off = off, off + nb_pad

        def run(st):
            arena.bytes_view(in_off, nb_pad)[:nb].copy_(flat, non_blocking=True)
            algo = ext.ALGO_NVLS if (arena.has_multicast and CONFIG.algo in ("auto", "nvls")) else ext.ALGO_TWOSHOT
            ext.allgather(arena.mesh, 0, algo, in_off, out_off, nb_pad, pick_blocks(nb_pad, 1, False), CONFIG.threads,
                          st.cuda_stream)
            res = arena.bytes_view(out_off, nb_pad * n).view(n, nb_pad)[:, :nb]
            out.view(n, -1).view(torch.uint8).copy_(res, non_blocking=True)

        _run_on_comm(plane, run, (flat, out))
        return out
    native = getattr(rt.control, "all_gather_cpu", None)
    if native is not None:
        return native(x.contiguous(), name)
    parts = rt.control.all_gather_bytes(x.contiguous().view(-1).view(torch.uint8).numpy().tobytes())
    for i, p in enumerate(parts):
        out[i].copy_(torch.frombuffer(bytearray(p), dtype=x.dtype).view(x.shape))
    return out


def all_gather_transform(x: torch.Tensor, f, name: str = "") -> torch.Tensor:
    """Gather every rank's `x`, apply `f(stacked [np, ...]) -> tensor shaped like x` ON RANK 0 and hand its result to everyone.
    Reference: `Peer::AllGatherTransform` = Gather -> f on the root -> Broadcast (srcs/cpp/src/session.cpp:162-181)."""
    rt = get_runtime()
    stacked = all_gather(x, name)
    out = torch.empty_like(x)
    if rt.rank == 0:
        y = f(stacked)
        if tuple(y.shape) != tuple(x.shape) or y.dtype != x.dtype:
            raise ValueError("all_gather_transform: f must return a tensor shaped and typed like its per-rank input")
        out.copy_(y)
    return broadcast_(out, 0, name)


def barrier() -> None:
    """Host barrier over the control plane (+ device flag barrier when a GPU plane exists)."""
    rt = get_runtime()
    if rt.size > 1 and rt._gpu is not None and rt._gpu._staging is not None:
        a, _, _ = rt.gpu.staging()
        _run_on_comm(rt.gpu, lambda s: rt.gpu.ext.barrier(a.mesh, 0, s.cuda_stream))
        torch.cuda.current_stream().synchronize()
        a.check()
    rt.control.barrier()


def consensus(x) -> bool:
    """True iff every worker passed identical bytes (tensor or bytes-like). Reference: `consensus` op / `BytesConsensus`."""
    rt = get_runtime()
    if isinstance(x, torch.Tensor):
        b = x.detach().cpu().contiguous().view(-1).view(torch.uint8).numpy().tobytes()
    else:
        b = bytes(x)
    return rt.control.consensus(b)
