"""PairAveragingOptimizer (AD-PSGD, asynchronous decentralised gossip).

Reference (TF only): srcs/python/kungfu/tensorflow/optimizers/async_sgd.py:13-142 — every step pick a random peer != self,
`request_variable(target)` pulls the peer's last *saved* fused model (first step: save own model, then barrier so the
first pull cannot miss), `v <- 0.5 (v + other)`, apply local gradients, `save_variable(new model)`. Legacy C++ ops with
`RandomSelector` / `RoundRobinSelector`: srcs/cpp/src/tensorflow/ops/cpu/peer_to_peer.cpp:8-524.

GPU: the model store is a TRIPLE-buffered region of the symmetric arena with version words (version v lives in buffer
v % 3, so a reader of v is only disturbed once its owner starts on v + 3 — two full steps later). `zero_grad()` picks the
step's target and starts `pull_tma_kernel` on the side stream: one elected thread per CTA drives the TMA engine and copies the
target's published model over NVLink into a private buffer while forward+backward run (the reference's "the pull overlaps
training", srcs/cpp/src/tensorflow/ops/cpu/peer_to_peer.cpp:212-254 AsyncModelAveraging); a copy that was torn anyway is
pulled again from the then-current version. After backward ONE local kernel averages, applies the optimizer step and
publishes. The target is never blocked. Without `zero_grad()` before backward the kernel reads the target over NVLink itself.
CPU: the native runtime's versioned blob store + P2P request (`rt.control.save / request`).
"""
from __future__ import annotations

import random
from typing import Optional

import torch

from ...ops import collective as coll
from .core import KungFuMixin, fusable_kind, wrap_optimizer


def get_random_peer(cluster_size: int, self_rank: int, rng: random.Random = None) -> int:
    """A uniformly drawn rank; a draw that hits `self_rank` moves on to the next rank, exactly like the reference's graph
    (async_sgd.py:71-75: `t = uniform(0, np); t == self ? (t + 1) % np : t`)."""
    if cluster_size <= 1:
        return int(self_rank)
    t = (rng or random).randrange(cluster_size)
    return (t + 1) % cluster_size if t == self_rank else t


class _PairAveraging(KungFuMixin):
    def _kf_init_pair(self, named_parameters, fuse_requests, fused_model_name, fused, peer_selection, seed, overlap=True):
        self._kf_setup(named_parameters, True)
        self._kf_fuse_requests = bool(fuse_requests)
        self._kf_model_name = fused_model_name or "kungfu::PairAveragingOptimizer::fused_model"
        self._kf_selection = peer_selection
        self._kf_rng = random.Random((seed if seed is not None else 0x5eed) * 1000003 + self._kf_rt.rank)
        self._kf_rr = 0
        kind = fusable_kind(self) if fused in (True, "auto") else None
        self._kf_kind = kind
        self._kf_published = False
        self._kf_next_target = None
        if self._kf_gpu:
            from ...parallel.engine import FlatEngine

            self._kf_engine = FlatEngine(self.param_groups, self._kf_named, mode="pair", opt_kind=kind or "sgd", average=False,
                                         overlap=overlap, name="pair", master_weights=kind is not None)

    def _kf_begin_step(self) -> None:
        """Choose this step's target and start pulling its model under forward+backward (GPU engine, after the first publish)."""
        eng = self._kf_engine
        if eng is None or self._kf_rt.size == 1 or not self._kf_published or self._kf_next_target is not None:
            return
        ver = getattr(self._kf_rt.control, "cluster_version", None)
        if getattr(self, "_kf_membership", None) != ((ver() if ver else 0), self._kf_rt.size):
            return  # membership changed: the next step() republishes first
        self._kf_next_target = self._kf_pick_peer()
        eng.begin_step(self._kf_next_target)

    def prepare_replay(self) -> None:
        """Before replaying a CUDA graph of a step that was captured with this optimizer: draw this replay's target and put
        it into the device word the captured pull / averaging kernels read."""
        eng = self._kf_engine
        if eng is not None and self._kf_rt.size > 1:
            eng.set_pair_target(self._kf_pick_peer())

    def _kf_first_publish(self) -> None:
        """First step of a membership (job start or right after a resize): every member publishes, then a barrier, so that no
        pull can find an empty store — joiners and survivors meet here together (reference: async_sgd.py:106-123)."""
        eng = self._kf_engine
        eng._sync_hyper()
        eng._ensure_state()
        eng.publish()
        coll.barrier()
        self._kf_published = True

    # ---- peer selection (reference: async_sgd.py:71-75, peer_to_peer.cpp selectors) -----------------
    def _kf_pick_peer(self) -> int:
        rt = self._kf_rt
        n, me = rt.size, rt.rank
        if n == 1:
            return me
        if self._kf_selection == "roundrobin":
            self._kf_rr += 1
            return (me + 1 + (self._kf_rr - 1) % (n - 1)) % n
        t = self._kf_rng.randrange(n - 1)
        return t if t < me else t + 1

    # ---- CPU model store ----------------------------------------------------------------------------
    def _kf_flat_cpu(self) -> torch.Tensor:
        return torch.cat([p.data.reshape(-1).float() for _, p in self._kf_named]).cpu()

    def _kf_unflat_cpu(self, flat: torch.Tensor, blend: float) -> None:
        off = 0
        for _, p in self._kf_named:
            n = p.numel()
            p.data.mul_(1.0 - blend).add_(flat[off:off + n].view_as(p).to(device=p.device, dtype=p.dtype), alpha=blend)
            off += n

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        rt = self._kf_rt
        eng = self._kf_engine
        if eng is not None:
            if self._kf_membership_changed() or not self._kf_published:
                self._kf_next_target = None
                self._kf_first_publish()
            target = self._kf_next_target if self._kf_next_target is not None else self._kf_pick_peer()
            self._kf_next_target = None
            if rt.size == 1:
                target = rt.rank
            if self._kf_kind is not None:
                eng.pair_step(target, do_step=True)
                eng.finish()
            else:
                eng.pair_step(target, do_step=False)
                eng.finish()
                self._kf_inner_step()
                eng.publish()
        else:
            ctl = rt.control
            if rt.size > 1 and hasattr(ctl, "save"):
                if self._kf_membership_changed() or not self._kf_published:
                    # first step of this membership (start of the job, or right after a resize): everybody publishes, then a
                    # barrier, so that no pull can find an empty store — joiners and survivors meet here together
                    ctl.save(self._kf_model_name, self._kf_flat_cpu())
                    ctl.barrier()
                    self._kf_published = True
                target = self._kf_pick_peer()
                mine = self._kf_flat_cpu()
                other = torch.empty_like(mine)
                try:
                    ok = ctl.request(target, self._kf_model_name, other)
                except RuntimeError as e:
                    # Asynchronous gossip has no rendezvous: the target may already have finished its last step and gone away (or be
                    # restarting). A missed pull is a local step, not an error (found by a test that failed once in ~10 suite runs:
                    # the first worker to finish exited while a slower one was still pulling from it).
                    ok = False
                    self._kf_missed_pulls = getattr(self, "_kf_missed_pulls", 0) + 1
                    if self._kf_missed_pulls == 1:
                        import warnings

                        warnings.warn(f"kungfu_b200: PairAveraging could not pull from rank {target} ({e}); continuing with a local step")
                if ok:
                    self._kf_unflat_cpu(other, 0.5)
                self._kf_inner_step()
                ctl.save(self._kf_model_name, self._kf_flat_cpu())
            else:
                self._kf_inner_step()
        self._kf_step_count += 1
        return loss


def PairAveragingOptimizer(optimizer: torch.optim.Optimizer, named_parameters=None, fuse_requests: bool = True,
                           fused_model_name: Optional[str] = None, fused="auto", peer_selection: str = "random",
                           seed: Optional[int] = None, overlap: bool = True):
    """Wrap `optimizer` with asynchronous pairwise model averaging (reference signature:
    `(optimizer, fuse_requests=True, fused_model_name=None, name=None)`)."""
    return wrap_optimizer(optimizer, _PairAveraging,
                          lambda o: o._kf_init_pair(named_parameters, fuse_requests, fused_model_name, fused, peer_selection, seed, overlap))
