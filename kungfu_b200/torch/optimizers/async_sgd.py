"""PairAveragingOptimizer (AD-PSGD, asynchronous decentralised gossip).

Reference (TF only): srcs/python/kungfu/tensorflow/optimizers/async_sgd.py:13-142 — every step pick a random peer != self,
`request_variable(target)` pulls the peer's last *saved* fused model (first step: save own model, then barrier so the
first pull cannot miss), `v <- 0.5 (v + other)`, apply local gradients, `save_variable(new model)`. Legacy C++ ops with
`RandomSelector` / `RoundRobinSelector`: srcs/cpp/src/tensorflow/ops/cpu/peer_to_peer.cpp:8-524.

GPU: the model store is a double-buffered region of the symmetric arena with version words; one `pair_avg_kernel` reads
the target's published model straight over NVLink, averages, applies the local optimizer step and publishes the result.
The target is never blocked; a reader that overlaps two publishes is counted as `torn` in the mesh status.
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
    def _kf_init_pair(self, named_parameters, fuse_requests, fused_model_name, fused, peer_selection, seed):
        self._kf_setup(named_parameters, True)
        self._kf_fuse_requests = bool(fuse_requests)
        self._kf_model_name = fused_model_name or "kungfu::PairAveragingOptimizer::fused_model"
        self._kf_selection = peer_selection
        self._kf_rng = random.Random((seed if seed is not None else 0x5eed) * 1000003 + self._kf_rt.rank)
        self._kf_rr = 0
        kind = fusable_kind(self) if fused in (True, "auto") else None
        self._kf_kind = kind
        self._kf_published = False
        if self._kf_gpu:
            from ...parallel.engine import FlatEngine

            self._kf_engine = FlatEngine(self.param_groups, self._kf_named, mode="pair", opt_kind=kind or "sgd", average=False,
                                         overlap=False, name="pair")

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
            if not self._kf_published:
                eng._sync_hyper()
                eng._ensure_state()
                eng.publish()
                coll.barrier()
                self._kf_published = True
            target = self._kf_pick_peer()
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
                ok = ctl.request(target, self._kf_model_name, other)
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
                           seed: Optional[int] = None):
    """Wrap `optimizer` with asynchronous pairwise model averaging (reference signature:
    `(optimizer, fuse_requests=True, fused_model_name=None, name=None)`)."""
    return wrap_optimizer(optimizer, _PairAveraging,
                          lambda o: o._kf_init_pair(named_parameters, fuse_requests, fused_model_name, fused, peer_selection, seed))
