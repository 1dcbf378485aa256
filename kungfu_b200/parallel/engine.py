"""FlatEngine — the GPU execution engine behind the distributed optimizers.

One engine owns a symmetric arena holding the flat gradient region (and, depending on the mode, the flat
parameter region, an averaging scratch region or the published-model double buffer), re-points
`param.data` / `param.grad` into it, and launches ONE fused kernel per bucket on a high-priority comm
stream as soon as backward has produced the bucket (post-accumulate-grad hooks), so communication +
optimizer math overlap the rest of backward. Everything is stream-ordered: no host synchronisation, no
per-step negotiation, CUDA-graph capturable (sequence numbers and hyper-parameters live on the device).

Replaces: `NCCLScheduler` / `LinearExecutor` / `NCCLThread` order agreement + blocking per-op stream sync
(srcs/cpp/src/nccl/scheduler.cpp:8-130, gpu_collective.cpp:146-155), the TF-stream spin-wait
(srcs/cpp/src/tensorflow/ops/gpu/collective.cpp:11-34), and the torch path's "all comm after backward"
(srcs/python/kungfu/torch/optimizers/sync_sgd.py:12-26).

Modes
  ssgd    fused reduce-scatter -> optimizer(shard) -> all-gather(params)       [K1b]
  grads   bucketed in-place all-reduce(avg) of gradients, then any torch optimizer  [K1]
  sma     all-reduce(model)/np -> blend -> local step                               [K8]
  pair    pull one peer's published model -> average -> local step -> publish       [K9]

Split-phase launches (default, KUNGFU_GATED=0 restores the single-kernel form): every collective body is bracketed by two
one-warp kernels — a gate (flag barrier: all ranks have reached this point of their comm streams) and a wait (all ranks'
body blocks have reported completion) — and the body itself never spins, so a rank that is early does not park the body's
CTAs on SMs its own backward still needs.

Overlap
  ssgd / grads : per bucket, from post-accumulate-grad hooks, under the rest of backward.
  sma          : the model average does not depend on this step's gradients: `begin_step()` (called by `zero_grad()`)
                 starts the out-of-place all-reduce params -> scratch on the comm stream under forward+backward; only the
                 local blend+step kernel runs after backward.
  pair         : `begin_step(target)` starts the TMA pull of the target's published model into a private buffer on the side
                 stream under forward+backward; the local average+step+publish kernel runs after backward.

`backend="nccl"` replaces OUR collective kernels by `torch.distributed` NCCL calls on the same streams, buckets, CUDA graph
and local optimizer kernels — the same-box baseline arm of the benchmarks ("a path that only calls NCCL is the baseline").
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from ..ops import collective as coll
from ..runtime import get_runtime
from ..utils import dtypes
from ..utils.trace import trace_scope
from .flat import Bucket, FlatLayout, group_params

HP = dict(lr=0, momentum=1, dampening=2, weight_decay=3, nesterov=4, grad_scale=5, first_step=6, beta1=7, beta2=8, eps=9,
          bias_c1=10, bias_c2=11, step=12, alpha=13, decoupled_wd=14)


def _view_like(flat: torch.Tensor, offset: int, p: torch.Tensor) -> torch.Tensor:
    """View of `flat[offset : offset + p.numel()]` with p's shape AND memory format (channels_last conv
    weights keep their NHWC strides, so cuDNN sees the layout it was given and autograd's gradient
    layout contract holds without extra copies)."""
    if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
        return flat.as_strided(p.shape, p.stride(), flat.storage_offset() + offset)
    if p.dim() == 5 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last_3d):
        return flat.as_strided(p.shape, p.stride(), flat.storage_offset() + offset)
    return flat[offset:offset + p.numel()].view(p.shape)


class _Set:
    """One homogeneous (param group, dtype) set with its regions and state."""

    def __init__(self, group_index: int, layout: FlatLayout) -> None:
        self.group_index = group_index
        self.layout = layout
        self.g_off = 0
        self.p_off = 0
        self.x_off = 0  # sma: avg scratch; pair: pub buffer 0
        self.x_off2 = 0  # pair: pub buffer 1
        self.x_off3 = 0  # pair: pub buffer 2
        self.ver_off = 0
        self.pulled: Optional[torch.Tensor] = None  # pair: private copy of the target's model (side-stream pull)
        self.pull_ok: Optional[torch.Tensor] = None  # pair: [version, torn count, torn flag, scratch]
        self.nccl_avg: Optional[torch.Tensor] = None  # sma over NCCL: the averaged model
        self.grads: Optional[torch.Tensor] = None
        self.params: Optional[torch.Tensor] = None
        self.hyper: Optional[torch.Tensor] = None
        self.hyper_cache: Optional[List[float]] = None
        self.master: List[Optional[torch.Tensor]] = []
        self.s1: List[Optional[torch.Tensor]] = []
        self.s2: List[Optional[torch.Tensor]] = []
        self.ready: List[int] = []
        self.done_counter: Optional[torch.Tensor] = None
        self.state_ready = False


class FlatEngine:
    def __init__(self, param_groups, named_parameters=None, mode: str = "ssgd", opt_kind: Optional[str] = None,
                 average: bool = True, overlap: bool = True, bucket_mb: Optional[float] = None, monitor: bool = False,
                 alpha: float = 0.1, algo: Optional[str] = None, name: str = "engine", backend: Optional[str] = None,
                 master_weights: bool = True) -> None:
        self.rt = get_runtime()
        self.plane = self.rt.gpu
        self.ext = self.plane.ext
        self.mode = mode
        self.opt_kind = opt_kind  # "sgd" | "adam" | None (grads mode)
        self.average = average
        self.backend = (backend or os.environ.get("KUNGFU_BACKEND", "kernels")).lower()  # "kernels" | "nccl" (baseline arm)
        if self.backend not in ("kernels", "nccl"):
            raise ValueError(f"kungfu_b200: unknown engine backend {self.backend!r}")
        self.gated = os.environ.get("KUNGFU_GATED", "1") != "0"
        if not master_weights:  # the wrapped torch optimizer updates the 16-bit parameters itself: no fp32 shadow to go stale
            self._MASTER_MODES = ()
        self.overlap_requested = overlap
        self.overlap = overlap and mode in ("ssgd", "grads")
        self._defer = False  # inside no_sync(): hooks do not launch
        self._early: Optional[str] = None  # sma / pair: what begin_step() has started for the current step
        self.monitor = monitor
        self.alpha = alpha
        self.algo = algo
        self.param_groups = param_groups
        self.nranks = self.rt.size
        self.rank = self.rt.rank
        mb = bucket_mb if bucket_mb is not None else float(os.environ.get("KUNGFU_BUCKET_MB", "8"))
        self.bucket_bytes = int(mb * (1 << 20))
        self.max_blocks = int(os.environ.get("KUNGFU_COMM_BLOCKS", "32"))
        self.threads = coll.CONFIG.threads
        named = list(named_parameters) if named_parameters is not None else None
        self.sets: List[_Set] = []
        whole = mode in ("pair", "sma")  # model averaging works on the whole model at once
        for s in group_params(param_groups, named):
            bb = (1 << 62) if whole else self.bucket_bytes
            tail = 0 if (whole or not self.overlap) else int(float(os.environ.get("KUNGFU_TAIL_BUCKET_MB", "1")) * (1 << 20))
            self.sets.append(_Set(s["group"], FlatLayout(s["items"], s["dtype"], bb, tail_bytes=tail)))
        if not self.sets:
            raise ValueError("kungfu_b200: optimizer has no trainable parameters")
        self._consensus_key = None  # checked lazily, once per membership (constructors never communicate)
        self._build_arena(name)
        self._adopt_tensors()
        self.sumsq = torch.zeros(4, dtype=torch.float32, device=self.plane.device)
        self._step_open = False  # a step's first kernel has been launched
        self._launched = [0 for _ in self.sets]  # next bucket to launch per set
        self._hooks = []
        self._slots: list = []  # (set, entry, param, arena view, its address) of every hooked parameter
        self._pending_event: Optional[torch.cuda.Event] = None
        self.steps = 0
        self.kernel_launches = 0
        self._full_state: Optional[Dict[str, torch.Tensor]] = None
        # Overlap mode "steals" gradients: `.grad` is None during backward, autograd hands us its freshly produced tensor
        # (no accumulate kernel), and each bucket moves its gradients into the arena with ONE multi-tensor copy right
        # before its fused kernel (instead of one `grad += new` launch per parameter).
        # The model-averaging modes reduce nothing per bucket, but they take the same hand-over: without it autograd accumulates
        # into the arena views with one `grad += new` kernel per parameter (BERT-large: 345 launches, 0.96 ms + 0.19 ms of zero fill
        # per step) instead of a few batched copies.
        self._handover = (mode in ("pair", "sma") and self.backend == "kernels" and os.environ.get("KUNGFU_STEAL_GRADS", "1") != "0"
                          and os.environ.get("KUNGFU_NATIVE_ENGINE", "1") != "0")
        self._steal = (self.overlap or self._handover) and os.environ.get("KUNGFU_STEAL_GRADS", "1") != "0"
        self._pending = [[[] for _ in s.layout.buckets] for s in self.sets]
        self._got = [[False] * len(s.layout.entries) for s in self.sets]
        self.plane.engines.append(self)
        # The native GradEngine (csrc/engine/grad_engine.cpp) owns the hot path of the bucketed modes: hand-over of stolen
        # gradients (one multi-tensor copy per bucket), bucket counting, in-order launches, stream/event wiring. Python keeps
        # what needs Python objects (re-pointing .grad), the NCCL baseline arm and the membership / state management.
        self._native = None
        self._held: List[torch.Tensor] = []  # stolen gradients whose bucket copy has not been enqueued yet (released in finish())
        if (self.mode in ("ssgd", "grads") or self._handover) and self.backend == "kernels" and os.environ.get("KUNGFU_NATIVE_ENGINE", "1") != "0":
            self._build_native()
        if self.overlap or self._handover:
            self._install_hooks()

    # ---------------------------------------------------------------------------------------------
    def _build_native(self) -> None:
        ext = self.ext
        algo = {"twoshot": ext.ALGO_TWOSHOT, "nvls": ext.ALGO_NVLS}.get((self.algo or coll.CONFIG.algo).lower(), ext.ALGO_AUTO)
        kind = ext.OPT_ADAM if self.opt_kind == "adam" else ext.OPT_SGD
        emode = ext.MODE_COPY if self._handover else (ext.MODE_SSGD if self.mode == "ssgd" else ext.MODE_GRADS)
        ne = ext.GradEngine(self.arena.mesh, emode, bool(self.average), bool(self.gated),
                            self.max_blocks, self.threads, algo)
        ne.set_comm_stream(self.plane.comm_stream.cuda_stream)
        for s in self.sets:
            lay = s.layout
            si = ne.add_set(s.g_off, s.p_off, dtypes.dtype_code(lay.dtype), s.hyper.data_ptr() if self.opt_kind is not None else 0, kind)
            for b in lay.buckets:
                ne.add_bucket(si, b.start * lay.esize, b.nbytes(lay.esize), b.numel)
            for e in lay.entries:
                ne.add_entry(si, e.bucket, e.offset * lay.esize, e.numel * lay.esize)
        if self.monitor:
            ne.set_monitor(self.sumsq.data_ptr())
        self._native = ne

    def _push_native_state(self) -> None:
        ptr = lambda t: t.data_ptr() if t is not None else 0
        for si, s in enumerate(self.sets):
            for b in s.layout.buckets:
                self._native.set_bucket_state(si, b.index, ptr(s.master[b.index]), ptr(s.s1[b.index]), ptr(s.s2[b.index]))

    def _check_consensus(self) -> None:
        key = self.plane.membership_key()
        if self._consensus_key == key:
            return
        self._consensus_key = key
        if self.nranks == 1:
            return
        import hashlib

        h = hashlib.sha1()
        for s in self.sets:
            h.update(s.layout.digest())
        h.update(f"{self.mode}:{self.opt_kind}:{self.bucket_bytes}".encode())
        if not self.rt.control.consensus(h.digest()):
            raise RuntimeError("kungfu_b200: ranks disagree on the parameter/bucket plan")

    def _build_arena(self, name: str) -> None:
        total = 1 << 20
        for s in self.sets:
            nb = (s.layout.nbytes + 4095) // 4096 * 4096
            n_regions = {"ssgd": 2, "grads": 1, "sma": 3, "pair": 5}[self.mode]
            total += n_regions * (nb + 4096) + 8192
        self.arena = self.plane.new_arena(total, name)
        for s in self.sets:
            nb = s.layout.nbytes
            s.g_off = self.arena.alloc(nb, align=4096)
            if self.mode in ("ssgd", "sma", "pair"):
                s.p_off = self.arena.alloc(nb, align=4096)
            if self.mode == "sma":
                s.x_off = self.arena.alloc(nb, align=4096)
            if self.mode == "pair":
                s.x_off = self.arena.alloc(nb, align=4096)
                s.x_off2 = self.arena.alloc(nb, align=4096)
                s.x_off3 = self.arena.alloc(nb, align=4096)
                s.ver_off = self.arena.alloc(256, align=256)
                s.done_counter = torch.zeros(4, dtype=torch.int32, device=self.plane.device)
                s.pull_ok = torch.zeros(4, dtype=torch.int32, device=self.plane.device)
                self._target_dev = torch.zeros(1, dtype=torch.int32, device=self.plane.device)
            s.hyper = torch.zeros(16, dtype=torch.float32, device=self.plane.device)

    def _adopt_tensors(self) -> None:
        """Re-point param.grad (and param.data where the mode needs it) into the arena."""
        for s in self.sets:
            lay = s.layout
            s.grads = self.arena.view(s.g_off, lay.total, lay.dtype)
            s.grads.zero_()
            if self.mode in ("ssgd", "sma", "pair"):
                s.params = self.arena.view(s.p_off, lay.total, lay.dtype)
                s.params.zero_()
            for e in lay.entries:
                p = e.param
                if s.params is not None:
                    pv = _view_like(s.params, e.offset, p)
                    pv.copy_(p.data)
                    p.data = pv
                gv = _view_like(s.grads, e.offset, p)
                if p.grad is not None:
                    gv.copy_(p.grad)
                p.grad = gv
            nb = len(lay.buckets)
            s.master = [None] * nb
            s.s1 = [None] * nb
            s.s2 = [None] * nb
            s.ready = [0] * nb
        torch.cuda.current_stream().synchronize()

    # ---------------------------------------------------------------------------------------------
    def _sync_hyper(self) -> None:
        """Mirror the torch optimizer's param_group hyper-parameters into device memory (only when changed)."""
        for s in self.sets:
            g = self.param_groups[s.group_index]
            v = [0.0] * 16
            v[HP["lr"]] = float(g.get("lr", 0.0))
            v[HP["weight_decay"]] = float(g.get("weight_decay", 0.0))
            v[HP["grad_scale"]] = (1.0 / self.nranks) if (self.average and self.mode == "ssgd" and self.backend == "kernels") else 1.0
            v[HP["alpha"]] = float(self.alpha)
            if self.opt_kind == "adam":
                b1, b2 = g.get("betas", (0.9, 0.999))
                v[HP["beta1"]], v[HP["beta2"]] = float(b1), float(b2)
                v[HP["eps"]] = float(g.get("eps", 1e-8))
                v[HP["decoupled_wd"]] = 1.0 if g.get("decoupled_weight_decay", g.get("_adamw", False)) else 0.0
            else:
                v[HP["momentum"]] = float(g.get("momentum", 0.0))
                v[HP["dampening"]] = float(g.get("dampening", 0.0))
                v[HP["nesterov"]] = 1.0 if g.get("nesterov", False) else 0.0
            idx = [0, 1, 2, 3, 4, 5, 7, 8, 9, 13, 14]  # configuration slots only: the device-side step counter, first-step
            # flag and Adam bias corrections (6, 10, 11, 12) belong to the kernels and survive resizes / lr changes
            if s.hyper_cache is None or any(v[i] != s.hyper_cache[i] for i in idx):
                s.hyper[idx] = torch.tensor([v[i] for i in idx], dtype=torch.float32, device=s.hyper.device)
                s.hyper_cache = v

    @property
    def _sharded(self) -> bool:
        """fp32 master weights / optimizer state are partitioned across ranks (fused S-SGD over our kernels only)."""
        return self.mode == "ssgd" and self.nranks > 1 and self.backend == "kernels"

    def _needs_state(self, s: _Set):
        g = self.param_groups[s.group_index]
        if self.opt_kind == "adam":
            return True, True
        return float(g.get("momentum", 0.0)) != 0.0, False

    def _ensure_state(self) -> None:
        """Lazily create fp32 master weights + optimizer state (after any user-side broadcast)."""
        for s in self.sets:
            if s.state_ready or self.mode == "grads":
                s.state_ready = True
                continue
            lay = s.layout
            need1, need2 = self._needs_state(s)
            dev = self.plane.device
            for b in lay.buckets:
                if self._sharded:
                    n = lay.shard_elems(b, self.nranks)
                    lo, hi = lay.shard_range(b, self.nranks, self.rank)
                else:
                    n = b.numel
                    lo, hi = b.start, b.start + b.numel
                if lay.dtype != torch.float32 and self.mode in self._MASTER_MODES:
                    m = torch.zeros(n, dtype=torch.float32, device=dev)
                    m[:hi - lo].copy_(s.params[lo:hi])
                    s.master[b.index] = m
                if need1:
                    s.s1[b.index] = torch.zeros(n, dtype=torch.float32, device=dev)
                if need2:
                    s.s2[b.index] = torch.zeros(n, dtype=torch.float32, device=dev)
            s.state_ready = True

    _MASTER_MODES = ("ssgd", "pair")  # modes whose kernels keep fp32 master weights for 16-bit parameters

    # ---------------------------------------------------------------------------------------------
    def no_sync(self):
        """Context manager for gradient accumulation: backward passes inside it only accumulate (no bucket is launched);
        the first backward outside it reduces and applies the accumulated gradients. Reference behaviour for comparison:
        the torch SynchronousSGDOptimizer reduces inside step(), so plain accumulation loops work there."""
        eng = self

        class _NoSync:
            def __enter__(self):
                eng._defer = True
                if eng._native is not None:
                    eng._native.set_deferred(True)

            def __exit__(self, *a):
                eng._defer = False
                if eng._native is not None:
                    eng._native.set_deferred(False)

        return _NoSync()

    def _adopt_unhooked_grads(self) -> None:
        """Gradients that did not come through autograd's accumulate hook (assigned by hand: `p.grad = t`): move them into their arena
        slot now, otherwise the flush would treat the parameter as unused and zero its slot."""
        if self._native is None or not self._steal or self._defer:
            return
        st = torch.cuda.current_stream().cuda_stream
        for si, ei, p, gv, gv_ptr in self._slots:
            g = p.grad
            if g is not None and g.data_ptr() != gv_ptr:
                gv.copy_(g)
                p.grad = gv
                if not self._step_open:
                    self._open_step()
                self._native.grad_ready(si, ei, 0, st)

    def _install_hooks(self) -> None:
        self._slots = []
        if self._native is not None:
            ne = self._native
            cur = torch.cuda.current_stream
            for si, s in enumerate(self.sets):
                for ei, e in enumerate(s.layout.entries):
                    gv = _view_like(s.grads, e.offset, e.param)

                    gv_ptr, gv_stride = gv.data_ptr(), gv.stride()
                    self._slots.append((si, ei, e.param, gv, gv_ptr))

                    def nhook(p, si=si, ei=ei, gv=gv, gv_ptr=gv_ptr, gv_stride=gv_stride):  # ONE native call per parameter
                        if self._defer:
                            return
                        if not self._step_open:
                            self._open_step()
                        g = p.grad
                        st = cur().cuda_stream
                        if g is None or g.data_ptr() == gv_ptr:
                            ne.grad_ready(si, ei, 0, st)  # already in its arena slot (or absent: flush() zero-fills)
                            return
                        if g.stride() == gv_stride and g.dtype == gv.dtype:
                            # a fresh autograd tensor with the slot's memory layout: the engine copies it with the rest of the
                            # bucket in ONE launch; keep it alive until then (the allocator must not hand its block to a later
                            # backward kernel before the copy has been enqueued)
                            self._held.append(g)
                            ne.grad_ready(si, ei, g.data_ptr(), st)
                        else:
                            gv.copy_(g)  # different layout (e.g. an expanded or transposed gradient): let torch convert it
                            ne.grad_ready(si, ei, 0, st)
                        p.grad = gv
                    self._hooks.append(e.param.register_post_accumulate_grad_hook(nhook))
            return
        for si, s in enumerate(self.sets):
            for ei, e in enumerate(s.layout.entries):
                gv = _view_like(s.grads, e.offset, e.param)

                def hook(p, si=si, bi=e.bucket, gv=gv, ei=ei):
                    if self._defer:
                        return  # gradient accumulation: autograd keeps adding into p.grad, nothing is launched
                    if self._got[si][ei]:
                        raise RuntimeError(
                            "kungfu_b200: a second backward() reached a parameter whose bucket was already reduced (and, on the "
                            "fused path, applied) in this step. Wrap all but the last backward of an accumulation loop in "
                            "`with optimizer.no_sync():`, or build the optimizer with overlap=False.")
                    g = p.grad
                    if g is not None and g.data_ptr() != gv.data_ptr():
                        # autograd produced a fresh tensor (.grad was None): queue it for the bucket's multi-tensor copy
                        self._pending[si][bi].append((gv, g))
                        p.grad = gv
                    self._got[si][ei] = True
                    self._on_grad_ready(si, bi)
                self._hooks.append(e.param.register_post_accumulate_grad_hook(hook))

    def remove_hooks(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks.clear()

    def _on_grad_ready(self, si: int, bi: int) -> None:
        s = self.sets[si]
        s.ready[bi] += 1
        # launch strictly in bucket order so that every rank issues the same kernel sequence
        while self._launched[si] < len(s.layout.buckets):
            b = s.layout.buckets[self._launched[si]]
            if s.ready[b.index] < len(b.entries):
                break
            self._launch_bucket(si, b, after_current=True)
            self._launched[si] += 1

    def _open_step(self) -> None:
        if self._step_open:
            return
        self._step_open = True
        self._check_consensus()
        self.arena.ensure()
        self._sync_hyper()
        self._ensure_state()
        if self._native is not None:
            self._push_native_state()
            self._native.open_step(torch.cuda.current_stream().cuda_stream)  # stream ordering + device-side hyper tick
            return
        cs = self.plane.comm_stream
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        cs.wait_event(ev)
        if self.opt_kind is not None:
            for s in self.sets:
                self.ext.hyper_tick(s.hyper.data_ptr(), cs.cuda_stream)
                self.kernel_launches += 1

    def _algo_for(self, s: _Set, nbytes: int) -> int:
        dt = dtypes.dtype_code(s.layout.dtype)
        return coll.pick_algo(self.arena, nbytes, dt, dtypes.SUM, inplace=True, force=self.algo)

    def _launch_bucket(self, si: int, b: Bucket, after_current: bool) -> None:
        s = self.sets[si]
        lay = s.layout
        self._open_step()
        cs = self.plane.comm_stream
        pend = self._pending[si][b.index]
        if pend:
            torch._foreach_copy_([a for a, _ in pend], [g for _, g in pend])
            pend.clear()
            after_current = True
        if after_current:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            cs.wait_event(ev)
        nbytes = b.nbytes(lay.esize)
        boff = b.start * lay.esize
        dt = dtypes.dtype_code(lay.dtype)
        flags = self.ext.FLAG_SUMSQ if self.monitor else 0
        sumsq = self.sumsq.data_ptr() if self.monitor else 0
        kind = self.ext.OPT_ADAM if self.opt_kind == "adam" else self.ext.OPT_SGD
        ptr = lambda t: t.data_ptr() if t is not None else 0
        mesh = self.arena.mesh
        gate = self.gated and self.nranks > 1 and self.backend == "kernels"

        def bracket(blocks, launch):
            """gate -> body (never spins) -> wait, or the single self-synchronising kernel"""
            if gate:
                self.ext.barrier(mesh, 0, cs.cuda_stream)
                launch(self.ext.FLAG_GATED)
                self.ext.wait_done(mesh, 0, 0, self.nranks * blocks, cs.cuda_stream)
                self.kernel_launches += 2
            else:
                launch(0)

        if self.mode == "ssgd":
            if self.nranks == 1 or self.backend == "nccl":
                if self.nranks > 1:
                    self._nccl_all_reduce(s.grads[b.start:b.start + b.numel], cs)
                self.ext.flat_opt(s.params.data_ptr() + boff, dt, s.grads.data_ptr() + boff, dt, ptr(s.master[b.index]),
                                  ptr(s.s1[b.index]), ptr(s.s2[b.index]), s.hyper.data_ptr(), kind, b.numel, cs.cuda_stream)
                if self.monitor:
                    self.ext.sumsq(s.grads.data_ptr() + boff, dt, b.numel, self.sumsq.data_ptr(), cs.cuda_stream)
                    self.ext.sumsq(s.grads.data_ptr() + boff, dt, b.numel, self.sumsq.data_ptr() + 4, cs.cuda_stream)
                    self.kernel_launches += 2
            else:
                algo = self._algo_for(s, nbytes)
                blocks = coll.pick_blocks(nbytes, self.nranks, True, self.max_blocks)
                bracket(blocks, lambda g: self.ext.ssgd_fused(
                    mesh, 0, algo, s.g_off + boff, s.p_off + boff, dt, nbytes, ptr(s.master[b.index]), ptr(s.s1[b.index]),
                    ptr(s.s2[b.index]), s.hyper.data_ptr(), kind, sumsq, flags | g, blocks, self.threads, cs.cuda_stream))
        elif self.mode == "grads":
            if self.nranks > 1 and self.backend == "nccl":
                self._nccl_all_reduce(s.grads[b.start:b.start + b.numel], cs)
                self.kernel_launches -= 1
            elif self.nranks > 1:
                algo = self._algo_for(s, nbytes)
                blocks = coll.pick_blocks(nbytes, self.nranks, True, self.max_blocks)
                scale = 1.0 / self.nranks if self.average else 1.0
                f = flags | (self.ext.FLAG_SCALE if self.average else 0)
                bracket(blocks, lambda g: self.ext.allreduce(mesh, 0, algo, s.g_off + boff, 0, nbytes, dt, dtypes.SUM, scale, sumsq,
                                                             f | g, blocks, self.threads, cs.cuda_stream))
            elif self.monitor:
                self.ext.sumsq(s.grads.data_ptr() + boff, dt, b.numel, self.sumsq.data_ptr(), cs.cuda_stream)
                self.ext.sumsq(s.grads.data_ptr() + boff, dt, b.numel, self.sumsq.data_ptr() + 4, cs.cuda_stream)
                self.kernel_launches += 1
            else:
                self.kernel_launches -= 1
        elif self.mode == "sma":
            grads = (s.grads.data_ptr() + boff) if self.opt_kind is not None else 0
            if self.nranks == 1:
                self.ext.flat_opt(s.params.data_ptr() + boff, dt, s.grads.data_ptr() + boff, dt, 0, ptr(s.s1[b.index]),
                                  ptr(s.s2[b.index]), s.hyper.data_ptr(), kind, b.numel, cs.cuda_stream)
            elif self._early == "sma" or self.backend == "nccl":
                # the average is already in the scratch region (begin_step started it under forward/backward): local blend + step
                if self._early != "sma":
                    self._sma_average(si, cs)
                avg = s.nccl_avg.data_ptr() + boff if self.backend == "nccl" else self.arena.view(s.x_off, lay.total, lay.dtype).data_ptr() + boff
                self.ext.sma_blend(s.params.data_ptr() + boff, avg, dt, nbytes, grads, dt, ptr(s.s1[b.index]), ptr(s.s2[b.index]),
                                   s.hyper.data_ptr(), kind, min(4 * self.plane.sm_count, max(1, nbytes // (16 * 512))), 512,
                                   cs.cuda_stream)
            else:
                algo = self._algo_for(s, nbytes)
                blocks = coll.pick_blocks(nbytes, self.nranks, True, self.max_blocks)
                self.ext.sma_fused(mesh, 0, algo, s.p_off + boff, s.x_off + boff, dt, nbytes, grads, dt,
                                   ptr(s.s1[b.index]), ptr(s.s2[b.index]), 0, dt, s.hyper.data_ptr(), kind, blocks,
                                   self.threads, cs.cuda_stream)
        else:
            raise AssertionError(self.mode)
        self.kernel_launches += 1

    # ---- NCCL baseline arm --------------------------------------------------------------------------
    def _nccl_all_reduce(self, t: torch.Tensor, stream: torch.cuda.Stream) -> None:
        """`ncclAllReduce` (average) of a flat arena view on the comm stream — what the reference's own NCCL path does per
        tensor (srcs/cpp/src/nccl/gpu_collective.cpp:146-166), here per bucket and stream-ordered."""
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("kungfu_b200: backend='nccl' needs torch.distributed (NCCL) to be initialised")
        with torch.cuda.stream(stream):
            dist.all_reduce(t, op=dist.ReduceOp.AVG if self.average or self.mode == "sma" else dist.ReduceOp.SUM)

    # ---- work that does not depend on this step's gradients -----------------------------------------
    def _sma_average(self, si: int, cs: torch.cuda.Stream) -> None:
        """scratch <- mean over ranks of the parameter region (out-of-place, split-phase, few CTAs: it has a whole
        forward+backward to hide under)."""
        s = self.sets[si]
        lay = s.layout
        dt = dtypes.dtype_code(lay.dtype)
        nbytes = lay.nbytes
        if self.backend == "nccl":
            if s.nccl_avg is None:
                s.nccl_avg = torch.empty_like(s.params)
            with torch.cuda.stream(cs):
                s.nccl_avg.copy_(s.params)
            self._nccl_all_reduce(s.nccl_avg, cs)
            return
        mesh = self.arena.mesh
        algo = self._algo_for(s, nbytes)
        blocks = coll.pick_blocks(nbytes, self.nranks, True, int(os.environ.get("KUNGFU_SMA_BLOCKS", "16")))
        self.ext.barrier(mesh, 0, cs.cuda_stream)
        self.ext.allreduce_to(mesh, 0, algo, s.p_off, s.x_off, nbytes, dt, dtypes.SUM, 1.0 / self.nranks, 0,
                              self.ext.FLAG_SCALE | self.ext.FLAG_GATED, blocks, self.threads, cs.cuda_stream)
        self.ext.wait_done(mesh, 0, 0, self.nranks * blocks, cs.cuda_stream)
        self.kernel_launches += 3

    def begin_step(self, target: Optional[int] = None) -> None:
        """Start, under the coming forward+backward, the part of the step that does not need its gradients
        (sma: the model average; pair: the pull of `target`'s published model). Called by `optimizer.zero_grad()`;
        without it `step()` does the same work after backward (correct, not overlapped)."""
        if self._early is not None or self.nranks == 1 or not self.overlap_requested:
            return
        if self.mode == "sma":
            self._open_step()
            cs = self.plane.comm_stream
            for si in range(len(self.sets)):
                self._sma_average(si, cs)
            self._early = "sma"
        elif self.mode == "pair" and target is not None and self._pair_ready:
            self._open_step()
            ss = self.plane.side_stream
            self._set_target(target)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            ss.wait_event(ev)
            pull_blocks = int(os.environ.get("KUNGFU_PULL_BLOCKS", "24"))
            for s in self.sets:
                lay = s.layout
                if s.pulled is None:
                    s.pulled = torch.empty(lay.nbytes, dtype=torch.uint8, device=self.plane.device)
                with torch.cuda.stream(ss):
                    s.pull_ok[2:].zero_()
                pubs = [s.x_off, s.x_off2, s.x_off3]
                nb = min(pull_blocks, (lay.nbytes + 32767) // 32768)
                for retry in (0, 1):  # the second attempt runs only if the first copy was torn
                    self.ext.pull_tma(self.arena.mesh, 1, 0, pubs, s.ver_off, s.pulled.data_ptr(), lay.nbytes, s.pull_ok.data_ptr(),
                                      retry, nb, ss.cuda_stream, self._target_dev.data_ptr())
                self.kernel_launches += 3
            self._early = "pair"

    _pair_ready = False  # the first publish of this membership has happened (nothing to pull before it)

    def _set_target(self, target: int) -> None:
        """The target rank lives in a device word the kernels read, so that one captured graph serves every target.
        Under capture the caller refreshes it before each replay (`set_pair_target`)."""
        if not torch.cuda.is_current_stream_capturing():
            self._target_dev.fill_(int(target))  # the value travels as a kernel argument: no host buffer to keep alive

    def set_pair_target(self, target: int) -> None:
        """Before replaying a captured pair-averaging step: choose whom this replay pulls from."""
        self._set_target(target)

    # ---------------------------------------------------------------------------------------------
    def pair_step(self, target: int, do_step: bool) -> None:
        """PairAveraging: v <- 0.5 (v + other) -> local step -> publish, one kernel per set (whole model). `other` is the
        private copy `begin_step(target)` pulled on the side stream under forward+backward; without it the kernel reads the
        target's published model over NVLink itself (same result, the transfer is then exposed)."""
        self._open_step()
        cur = torch.cuda.current_stream()
        if self._handover:
            self._adopt_unhooked_grads()
            self._native.flush(cur.cuda_stream, True)  # gradients still held by autograd's tensors -> arena (zeros for unused parameters)
        early = self._early == "pair"
        tick = torch.cuda.Event()  # the hyper-parameter tick of _open_step went to the comm stream
        tick.record(self.plane.comm_stream)
        cur.wait_event(tick)
        if early:
            done = torch.cuda.Event()
            done.record(self.plane.side_stream)
            cur.wait_event(done)
        else:
            self._set_target(target)
        kind = self.ext.OPT_ADAM if self.opt_kind == "adam" else self.ext.OPT_SGD
        ptr = lambda t: t.data_ptr() if t is not None else 0
        for s in self.sets:
            lay = s.layout
            dt = dtypes.dtype_code(lay.dtype)
            blocks = coll.pick_blocks(lay.nbytes, 1, False, 148)  # after backward, nothing else runs: use every SM
            self.ext.pair_avg(self.arena.mesh, 1, 0, [s.x_off, s.x_off2, s.x_off3], s.ver_off, s.params.data_ptr(), dt, lay.nbytes,
                              s.grads.data_ptr(), dt, ptr(s.s1[0]), ptr(s.s2[0]), s.hyper.data_ptr(), kind, 1 if do_step else 0,
                              s.done_counter.data_ptr(), s.pulled.data_ptr() if early else 0, s.pull_ok.data_ptr() if early else 0,
                              blocks, self.threads, cur.cuda_stream, self._target_dev.data_ptr(), ptr(s.master[0]))
            self.kernel_launches += 1

    def publish(self) -> None:
        self._check_consensus()
        self.arena.ensure()
        cur = torch.cuda.current_stream()
        for s in self.sets:
            lay = s.layout
            blocks = coll.pick_blocks(lay.nbytes, 1, False, max(self.max_blocks, 64))
            self.ext.publish(self.arena.mesh, 1, [s.x_off, s.x_off2, s.x_off3], s.ver_off, s.params.data_ptr(), lay.nbytes,
                             s.done_counter.data_ptr(), blocks, self.threads, cur.cuda_stream)
            self.kernel_launches += 1
        self._pair_ready = True

    def torn_reads(self) -> int:
        """PairAveraging: pulls whose source buffer was handed back to its owner while being read (after the retry).
        Host-synchronising; for tests and end-of-run reports."""
        n = int(self.arena.mesh.status()[5])
        for s in self.sets:
            if s.pull_ok is not None:
                n += int(s.pull_ok[1].item())
        return n

    # ---------------------------------------------------------------------------------------------
    def flush(self) -> None:
        """Launch every bucket that the hooks have not launched yet (unused params, overlap disabled)."""
        if self._native is not None:
            if not self._step_open:
                self._open_step()
            self._adopt_unhooked_grads()
            self._native.flush(torch.cuda.current_stream().cuda_stream, bool(self._steal))
            if not self._handover:
                return
        elif self._steal:
            for si, s in enumerate(self.sets):
                for ei, e in enumerate(s.layout.entries):
                    if not self._got[si][ei] and self._launched[si] <= e.bucket:
                        # a parameter that received no gradient this step contributes zeros
                        _view_like(s.grads, e.offset, e.param).zero_()
        for si, s in enumerate(self.sets):
            first = True
            while self._launched[si] < len(s.layout.buckets):
                b = s.layout.buckets[self._launched[si]]
                self._launch_bucket(si, b, after_current=first)
                first = False
                self._launched[si] += 1

    def finish(self) -> None:
        """Make the current stream wait for the comm stream and reset the per-step bookkeeping."""
        if self._native is not None:
            status = self._native.finish(torch.cuda.current_stream().cuda_stream)
            self._held.clear()
            if not self._handover:
                self._step_open = False
                self._early = None
                self.steps += 1
                if status != 0:
                    self.arena.check()
                return
        done = torch.cuda.Event()
        done.record(self.plane.comm_stream)
        torch.cuda.current_stream().wait_event(done)
        self._step_open = False
        self._early = None
        if self.nranks > 1 and self.arena.mesh.status()[0] != 0:
            self.arena.check()  # a device-side wait timed out (host-mapped word: no synchronisation needed to read it)
        for si, s in enumerate(self.sets):
            self._launched[si] = 0
            for i in range(len(s.ready)):
                s.ready[i] = 0
            for i in range(len(self._got[si])):
                self._got[si][i] = False
        self.steps += 1

    def step_buckets(self) -> None:
        with trace_scope(f"engine.step[{self.mode}]"):
            self.flush()
            self.finish()

    def zero_grad(self) -> None:
        if self._steal:
            for s in self.sets:
                for e in s.layout.entries:
                    e.param.grad = None  # nothing to clear: the next backward hands over fresh gradient tensors
            return
        for s in self.sets:
            s.grads.zero_()

    def check(self) -> None:
        self.arena.check()

    # ---- elastic membership ------------------------------------------------------------------------
    def prepare_resize(self) -> None:
        """Called on every CURRENT member right before the membership changes: un-shard the fp32 master weights and
        optimizer state (they are partitioned by the old world size) while the old peer mappings still exist."""
        torch.cuda.synchronize(self.plane.device_index)
        self._full_state = None
        if all(s.state_ready for s in self.sets) and self.mode in ("ssgd", "sma", "pair") and self.opt_kind is not None:
            self._full_state = self.gather_state()

    def finish_resize(self) -> None:
        """Called after the arenas have been reconnected to the new membership."""
        self.rank, self.nranks = self.rt.rank, self.rt.size
        for s in self.sets:
            s.hyper_cache = None
            hp_step = s.hyper.clone()
            s.state_ready = False
            s._saved_hyper = hp_step  # type: ignore[attr-defined]
        self._step_open = False
        self._early = None
        self._pair_ready = False
        self._launched = [0 for _ in self.sets]

    def prepare_recover(self) -> None:
        """Called on every SURVIVOR after a peer died (`kungfu_b200.python.recover`), before the membership changes. Unlike
        `prepare_resize` nothing collective can run (the dead rank would never answer): the fp32 shards stay where they are,
        tagged with the world they were cut for, and `rebuild_state_after_loss` reassembles them once the survivors are
        reconnected. Device-side waits that ran into their time-out have poisoned the status word: clear it."""
        try:
            torch.cuda.synchronize(self.plane.device_index)
        except Exception:
            pass
        for a in list(self.plane._arenas):
            try:
                a.mesh.clear_status()
            except Exception:
                pass
        self._full_state = None
        self._held = []
        ready = all(s.state_ready for s in self.sets) and self.mode in ("ssgd", "sma", "pair") and self.opt_kind is not None
        self._lost = (self.rank, self.nranks, self._sharded) if ready else None

    def rebuild_state_after_loss(self) -> None:
        """COLLECTIVE over the survivors, after the arenas were reconnected: every survivor contributes the shards it owned in
        the OLD world; the ranges the dead ranks owned are refilled from the (replicated) low-precision parameters for the fp32
        master weights and with zeros for momentum / Adam moments — that slice of the optimizer history is gone, the model is
        not. The result is the un-sharded state `sync_state_from_root` distributes."""
        lost = getattr(self, "_lost", None)
        self._lost = None
        # who has what: agreed on the host plane (one small message; the survivors' new session is already up)
        mine = b"-" if lost is None else f"{lost[0]}/{lost[1]}/{int(lost[2])}".encode()
        views = self.rt.control.all_gather_bytes(mine) if self.nranks > 1 else [mine]
        if lost is None or any(v == b"-" for v in views) or len({v.split(b"/", 1)[1] for v in views}) != 1:
            return  # somebody never stepped (or the survivors disagree about the old world): lazily initialised state is right
        old_rank, old_n, was_sharded = lost
        dev = self.plane.device
        full: Dict[str, torch.Tensor] = {}
        owners = [int(v.split(b"/")[0]) for v in views]
        for si, s in enumerate(self.sets):
            lay = s.layout
            for key, shards in (("master", s.master), ("s1", s.s1), ("s2", s.s2)):
                if not any(t is not None for t in shards):
                    continue
                parts = []
                for b in lay.buckets:
                    t = shards[b.index]
                    if t is None:
                        continue
                    if not was_sharded:
                        parts.append(t[:b.numel].clone())
                        continue
                    per = lay.shard_elems(b, old_n)
                    if key == "master":
                        whole = torch.zeros(per * old_n, dtype=torch.float32, device=dev)
                        whole[:b.numel].copy_(s.params[b.start:b.start + b.numel])
                    else:
                        whole = torch.zeros(per * old_n, dtype=torch.float32, device=dev)
                    got = coll.all_gather(t.contiguous()).view(len(owners), -1) if self.nranks > 1 else t.view(1, -1)
                    for i, r in enumerate(owners):
                        whole[r * per:(r + 1) * per].copy_(got[i][:per])
                    parts.append(whole[:b.numel])
                full[f"set{si}.{key}"] = torch.cat(parts)
        self._full_state = full

    def sync_state_from_root(self, root: int = 0) -> None:
        """Collective over the NEW membership (survivors and newcomers): rank `root` broadcasts the full fp32 state,
        every rank keeps the shard it now owns. Newcomers call it with no local state."""
        if self.mode == "grads" or self.opt_kind is None:
            for s in self.sets:
                s.state_ready = self.mode == "grads"
            return
        dev = self.plane.device
        for si, s in enumerate(self.sets):
            lay = s.layout
            need1, need2 = self._needs_state(s)
            keys = [("master", lay.dtype != torch.float32 and self.mode in self._MASTER_MODES), ("s1", need1), ("s2", need2)]
            # the device-side step counter / bias corrections travel too
            coll.broadcast_(s.hyper, root)
            s.hyper_cache = None
            for key, needed in keys:
                if not needed:
                    continue
                full = None
                if self._full_state is not None:
                    full = self._full_state.get(f"set{si}.{key}")
                have = torch.tensor([1 if full is not None else 0], dtype=torch.int32, device=dev)
                coll.broadcast_(have, root)
                if int(have.item()) == 0:
                    continue  # the root never stepped: lazily initialised state is right for everybody
                if full is None or full.numel() != sum(b.numel for b in lay.buckets):
                    full = torch.zeros(sum(b.numel for b in lay.buckets), dtype=torch.float32, device=dev)
                full = full.contiguous()
                coll.broadcast_(full, root)
                self._install_full(s, key, full)
            # anything not received is created lazily (zeros / from the broadcast parameters)
            s.state_ready = False
        self._ensure_state_keep_existing()
        self._full_state = None

    def _install_full(self, s: _Set, key: str, full: torch.Tensor) -> None:
        """Cut the un-sharded fp32 state `full` (all buckets, concatenated) into this rank's per-bucket shards."""
        lay = s.layout
        dev = self.plane.device
        off = 0
        shards = getattr(s, key)
        for b in lay.buckets:
            if self._sharded:
                n = lay.shard_elems(b, self.nranks)
                lo, hi = lay.shard_range(b, self.nranks, self.rank)
            else:
                n = b.numel
                lo, hi = b.start, b.start + b.numel
            t = torch.zeros(n, dtype=torch.float32, device=dev)
            t[:hi - lo].copy_(full[off + (lo - b.start):off + (hi - b.start)])
            shards[b.index] = t
            off += b.numel

    # ---- checkpointing -----------------------------------------------------------------------------
    def state_dict(self) -> Dict[str, object]:
        """COLLECTIVE when the state is sharded (every rank must call it): the un-sharded fp32 master weights and optimizer
        state plus the device-side step counters, on the host. The reference has no optimizer checkpoint at all."""
        torch.cuda.current_stream().synchronize()
        full = self.gather_state() if all(s.state_ready for s in self.sets) else {}
        return {"mode": self.mode, "opt_kind": self.opt_kind, "layout": [s.layout.digest().hex() for s in self.sets],
                "state": {k: v.detach().cpu() for k, v in full.items()}, "hyper": [s.hyper.detach().cpu() for s in self.sets]}

    def load_state_dict(self, sd: Dict[str, object]) -> None:
        """Local (every rank reads the same file): re-shard the saved state for the CURRENT world size and restore the
        device-side step counters. Call after the model parameters have been loaded."""
        if sd.get("mode") != self.mode or sd.get("opt_kind") != self.opt_kind:
            raise ValueError(f"kungfu_b200: checkpoint is for engine {sd.get('mode')}/{sd.get('opt_kind')}, this one is {self.mode}/{self.opt_kind}")
        if list(sd.get("layout", [])) != [s.layout.digest().hex() for s in self.sets]:
            raise ValueError("kungfu_b200: checkpoint was written for a different parameter/bucket layout")
        dev = self.plane.device
        state = sd.get("state", {})
        for si, s in enumerate(self.sets):
            for key in ("master", "s1", "s2"):
                full = state.get(f"set{si}.{key}")
                if full is not None:
                    self._install_full(s, key, full.to(dev, torch.float32))
            s.hyper.copy_(sd["hyper"][si].to(dev))
            s.hyper_cache = None
            s.state_ready = False
        self._ensure_state_keep_existing()  # whatever the checkpoint did not contain (a run that never stepped)
        self._step_open = False
        self._early = None
        self._pair_ready = False
        torch.cuda.current_stream().synchronize()

    def _ensure_state_keep_existing(self) -> None:
        for s in self.sets:
            lay = s.layout
            need1, need2 = self._needs_state(s)
            dev = self.plane.device
            for b in lay.buckets:
                if self._sharded:
                    n = lay.shard_elems(b, self.nranks)
                    lo, hi = lay.shard_range(b, self.nranks, self.rank)
                else:
                    n = b.numel
                    lo, hi = b.start, b.start + b.numel
                ok = lambda t: t is not None and t.numel() == n
                if lay.dtype != torch.float32 and self.mode in self._MASTER_MODES and not ok(s.master[b.index]):
                    m = torch.zeros(n, dtype=torch.float32, device=dev)
                    m[:hi - lo].copy_(s.params[lo:hi])
                    s.master[b.index] = m
                if need1 and not ok(s.s1[b.index]):
                    s.s1[b.index] = torch.zeros(n, dtype=torch.float32, device=dev)
                if need2 and not ok(s.s2[b.index]):
                    s.s2[b.index] = torch.zeros(n, dtype=torch.float32, device=dev)
            s.state_ready = True

    # ---- monitoring --------------------------------------------------------------------------------
    def read_sumsq(self, reset: bool = True) -> torch.Tensor:
        """[sum_r |g_r|^2, |g_avg|^2] summed over ranks (device tensor, stream-ordered)."""
        out = self.sumsq[:2].clone()
        if self.nranks > 1:
            coll.all_reduce_(out, "sum")
        if reset:
            self.sumsq.zero_()
        return out

    def gather_state(self) -> Dict[str, torch.Tensor]:
        """Full (un-sharded) fp32 optimizer state for checkpointing: {master, s1, s2} flat per set."""
        out: Dict[str, torch.Tensor] = {}
        for si, s in enumerate(self.sets):
            for key, shards in (("master", s.master), ("s1", s.s1), ("s2", s.s2)):
                if not any(t is not None for t in shards):
                    continue
                parts = []
                for b in s.layout.buckets:
                    t = shards[b.index]
                    if t is None:
                        continue
                    if self._sharded:
                        g = coll.all_gather(t).view(-1)[:b.numel]
                    else:
                        g = t[:b.numel]
                    parts.append(g)
                out[f"set{si}.{key}"] = torch.cat(parts)
        return out
