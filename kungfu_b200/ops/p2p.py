"""Peer-to-peer model store ops.

Reference: `save_variable(s)` / `request_variable(_with_template)` (srcs/python/kungfu/tensorflow/ops/p2p.py:4-26,
local.py; kernels cpu/local.cpp:29-43, cpu/p2p_new.cpp:39-57). A request for an unknown name/version reports failure
instead of raising (rchannel/handler/p2p.go:102-120).
CPU tensors use the native runtime's blob store; CUDA tensors use a published double buffer in the symmetric arena
that peers read directly over NVLink (`publish` / `pull` kernels).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import os

import torch

from ..runtime import get_runtime

_gpu_slots: Dict[str, dict] = {}


def _gpu_slot(name: str, nbytes: int) -> dict:
    slot = _gpu_slots.get(name)
    if slot is None:
        rt = get_runtime()
        plane = rt.gpu
        pad = (nbytes + 15) // 16 * 16
        arena = plane.new_arena(2 * pad + (1 << 16), f"p2p:{name}", want_multicast=False)
        slot = {"arena": arena, "pub0": arena.alloc(pad, 4096), "pub1": arena.alloc(pad, 4096), "ver": arena.alloc(256, 256),
                "nbytes": pad, "done": torch.zeros(4, dtype=torch.int32, device=plane.device),
                "ok": torch.zeros(4, dtype=torch.int32, device=plane.device),
                "stage": torch.zeros(pad, dtype=torch.uint8, device=plane.device)}
        _gpu_slots[name] = slot
    slot["arena"].ensure()
    if slot["nbytes"] < nbytes:
        raise ValueError(f"p2p slot {name!r} was created with {slot['nbytes']} bytes < {nbytes}")
    return slot


def save_variable(t: torch.Tensor, name: str, version: Optional[int] = None) -> None:
    """Publish `t` under `name` so that peers can `request_variable` it. Collective on first use for CUDA tensors
    (every rank must create the slot)."""
    rt = get_runtime()
    if t.is_cuda and rt.gpu.spans_hosts:
        t = t.detach().to("cpu")  # peers on other hosts cannot map this GPU: the blob goes through the host store
    if t.is_cuda:
        nb = t.numel() * t.element_size()
        s = _gpu_slot(name, nb)
        st = torch.cuda.current_stream()
        s["stage"][:nb].copy_(t.detach().contiguous().view(-1).view(torch.uint8), non_blocking=True)
        rt.gpu.ext.publish(s["arena"].mesh, 1, [s["pub0"], s["pub1"]], s["ver"], s["stage"].data_ptr(), s["nbytes"], s["done"].data_ptr(),
                           16, 512, st.cuda_stream)
        return
    f = getattr(rt.control, "save", None)
    if f is None:
        _local_store[(name, version)] = t.detach().clone()
        return
    f(name, t.detach(), None if version is None else str(version))


_local_store: Dict[Tuple[str, Optional[int]], torch.Tensor] = {}


_TMA_MIN_BYTES = int(os.environ.get("KUNGFU_PULL_TMA_MIN_BYTES", str(256 << 10)))


def request_variable(target: int, name: str, shape=None, dtype=None, version: Optional[int] = None,
                     template: Optional[torch.Tensor] = None, device=None) -> Tuple[bool, torch.Tensor]:
    """Pull the blob `name` last saved by rank `target`. Returns (ok, tensor); ok is False (tensor untouched/zero)
    when the target has nothing under that name/version."""
    rt = get_runtime()
    if template is not None:
        shape, dtype, device = template.shape, template.dtype, template.device
    out = torch.zeros(tuple(shape), dtype=dtype, device=device or "cpu")
    if out.is_cuda and rt.gpu.spans_hosts:
        ok, host = request_variable(target, name, shape, dtype, version, device="cpu")
        if ok:
            out.copy_(host)
        return ok, out
    if out.is_cuda:
        nb = out.numel() * out.element_size()
        s = _gpu_slot(name, nb)
        st = torch.cuda.current_stream()
        tmp = torch.empty(s["nbytes"], dtype=torch.uint8, device=out.device)
        if s["nbytes"] >= _TMA_MIN_BYTES:
            # large blobs: TMA bulk copies (cp.async.bulk through shared memory), one CTA per 32 KiB chunk up to one per SM
            blocks = min(148, (s["nbytes"] + 32767) // 32768)
            rt.gpu.ext.pull_tma(s["arena"].mesh, 1, int(target), [s["pub0"], s["pub1"]], s["ver"], tmp.data_ptr(), s["nbytes"],
                                s["ok"].data_ptr(), 0, blocks, st.cuda_stream)
        else:
            rt.gpu.ext.pull(s["arena"].mesh, 1, int(target), [s["pub0"], s["pub1"]], s["ver"], tmp.data_ptr(), s["nbytes"],
                            s["ok"].data_ptr(), 0, 16, 512, st.cuda_stream)
        ok = int(s["ok"][0].item()) != 0
        if ok:
            out.view(-1).view(torch.uint8).copy_(tmp[:nb])
        return ok, out
    f = getattr(rt.control, "request", None)
    if f is None:
        t = _local_store.get((name, version))
        if t is None or target != rt.rank:
            return False, out
        out.copy_(t)
        return True, out
    ok = f(int(target), name, out, None if version is None else str(version))
    return bool(ok), out


def request_variable_with_template(target: int, template: torch.Tensor, name: str, version: Optional[int] = None):
    return request_variable(target, name, version=version, template=template)


def save_variables(tensors, names, version: Optional[int] = None) -> None:
    for t, n in zip(tensors, names):
        save_variable(t, n, version)


# ---- legacy fused-model ops (reference: ModelAveraging / AsyncModelAveraging / SaveModel / RequestModel / AsyncRequestModel with
# RandomSelector / RoundRobinSelector, srcs/cpp/src/tensorflow/ops/cpu/peer_to_peer.cpp:8-524) -----------------------------------------
class RoundRobinSelector:
    """Next peer in rank order, skipping oneself (peer_to_peer.cpp `RoundRobinSelector`)."""

    def __init__(self, size: int, rank: int) -> None:
        self.peers = [r for r in range(size) if r != rank] or [rank]
        self.i = 0

    def next(self) -> int:
        p = self.peers[self.i % len(self.peers)]
        self.i += 1
        return p


class RandomSelector:
    """A uniformly drawn peer other than oneself (peer_to_peer.cpp `RandomSelector`)."""

    def __init__(self, size: int, rank: int, seed: Optional[int] = None) -> None:
        import random

        self.peers = [r for r in range(size) if r != rank] or [rank]
        self.rng = random.Random(seed if seed is not None else 0xA11CE + rank)

    def next(self) -> int:
        return self.rng.choice(self.peers)


def _fuse(vs):
    return torch.cat([v.detach().reshape(-1) for v in vs])


def _blend_in(vs, flat: torch.Tensor) -> None:
    off = 0
    with torch.no_grad():
        for v in vs:
            n = v.numel()
            v.mul_(0.5).add_(flat[off:off + n].view_as(v).to(v.dtype), alpha=0.5)
            off += n


class ModelAveraging:
    """`model_averaging(variables)`: pull the fused model of a selected peer NOW and average it in (`v <- 0.5 (v + other)`);
    the requester waits for the transfer. Call `save_model` after each local update so that peers find something to pull."""

    def __init__(self, variables, strategy: str = "random", name: str = "kungfu::legacy_model", seed: Optional[int] = None) -> None:
        rt = get_runtime()
        self.vs = list(variables)
        self.name = name
        self.sel = RoundRobinSelector(rt.size, rt.rank) if strategy == "roundrobin" else RandomSelector(rt.size, rt.rank, seed)

    def save(self) -> None:
        save_variable(_fuse(self.vs), self.name)

    def __call__(self) -> bool:
        if get_runtime().size == 1:
            return False
        ok, flat = request_variable_with_template(self.sel.next(), _fuse(self.vs), self.name)
        if ok:
            _blend_in(self.vs, flat)
        return bool(ok)


class AsyncModelAveraging(ModelAveraging):
    """`async_model_averaging(variables)`: double-buffered PREFETCH — every call averages in the model that was fetched while the
    previous training step ran and starts fetching the next peer's model; the caller never waits for a transfer (reference:
    peer_to_peer.cpp:212-254, whose prefetch callback races with the averaging — `is_requesting_` is read unlocked; here the
    hand-over is a stream event on GPUs and a joined future on the host)."""

    def __init__(self, variables, strategy: str = "random", name: str = "kungfu::legacy_model", seed: Optional[int] = None) -> None:
        super().__init__(variables, strategy, name, seed)
        self._pending = None  # (ok_or_future, buffer, event)
        self._pool = None

    def _start_prefetch(self) -> None:
        target = self.sel.next()
        template = _fuse(self.vs)
        if template.is_cuda:
            rt = get_runtime()
            side = rt.gpu.side_stream
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                ok, flat = request_variable_with_template(target, template, self.name)
                ev = torch.cuda.Event()
                ev.record(side)
            self._pending = (ok, flat, ev)
        else:
            if self._pool is None:
                from concurrent.futures import ThreadPoolExecutor

                self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="kf-prefetch")
            self._pending = (self._pool.submit(request_variable_with_template, target, template, self.name), None, None)

    def __call__(self) -> bool:
        if get_runtime().size == 1:
            return False
        averaged = False
        if self._pending is not None:
            ok, flat, ev = self._pending
            if ev is None:  # host: the future holds (ok, tensor)
                ok, flat = ok.result()
            else:
                torch.cuda.current_stream().wait_event(ev)
            if ok:
                _blend_in(self.vs, flat)
                averaged = True
        self._start_prefetch()
        return averaged


def save_model_legacy(variables, name: str = "kungfu::legacy_model") -> None:
    """`SaveModel`: publish the fused model under `name`."""
    save_variable(_fuse(list(variables)), name)
