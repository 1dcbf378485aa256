"""MonitorGradientNoiseScaleOptimizer: S-SGD + online gradient-noise-scale estimation.

Reference (TF only): srcs/python/kungfu/tensorflow/optimizers/grad_noise_scale.py:11-88 + ops/monitor.py:6-17 + kernel
`KungfuNoiseScale` (cpu/collective.cpp:212-258): every `monitor_interval` steps compute G_small = |local grads|^2,
G_big = |averaged grads|^2 (two extra full passes over the gradients), biased estimators, two EMAs, ratio.

GPU: both squared norms are side reductions INSIDE the fused all-reduce / S-SGD kernels (flag kArSumSq) — zero extra
passes over the gradients — and the estimators + EMAs are finalised by `gns_finalize_kernel` on the device. G_small is
the mean over workers of |g_r|^2 (same expectation as the reference's per-worker value, lower variance).
"""
from __future__ import annotations

from typing import Optional

import torch

from ...ops import collective as coll
from ...ops.monitor import NoiseScale
from .core import wrap_optimizer
from .sync_sgd import _SynchronousSGD


def _publish_gauge(v: float) -> None:
    from ...monitor import set_gauge

    set_gauge("kungfu_gradient_noise_scale", v)


class _MonitorGNS(_SynchronousSGD):
    def _kf_init_gns(self, named_parameters, device_batch_size, monitor_interval, alpha, fused, overlap, bucket_mb, algo):
        self._kf_init_ssgd(named_parameters, "avg", fused, overlap, bucket_mb, True, algo)
        self._kf_bs = float(device_batch_size)
        self._kf_interval = int(monitor_interval)
        self._kf_alpha = float(alpha)
        self._kf_cpu_ns = NoiseScale(alpha)
        self.noise_scale_value = 0.0
        from ... import variables as kv

        if self._kf_engine is not None:
            dev = self._kf_engine.plane.device
            self._kf_gns_state = torch.zeros(4, dtype=torch.float32, device=dev)
            self._kf_gns_out = torch.zeros(4, dtype=torch.float32, device=dev)
            # reference: the optimizer owns the global variable `kungfu_gradient_noise_scale`; here it is the device scalar the
            # finalize kernel writes (reading it with eval_gradient_noise_scale() synchronises, nothing else does)
            self._kf_gns_var = kv.bind_global_variable(kv.GraphKeys.GRADIENT_NOISE_SCALE, self._kf_gns_out[0])
        else:
            self._kf_gns_var = kv.bind_global_variable(kv.GraphKeys.GRADIENT_NOISE_SCALE, torch.zeros((), dtype=torch.float32))

    @torch.no_grad()
    def step(self, closure=None):
        rt = self._kf_rt
        n = rt.size
        eng = self._kf_engine
        if eng is None:
            # CPU path: explicit norms like the reference
            local_sq = sum(float((p.grad.detach().float() ** 2).sum()) for _, p in self._kf_named if p.grad is not None)
            loss = _SynchronousSGD.step(self, closure)  # averages grads in place, then steps
            # NB: after step() p.grad holds the averaged gradient
            big_sq = sum(float((p.grad.detach().float() ** 2).sum()) for _, p in self._kf_named if p.grad is not None)
            if n > 1:
                t = torch.tensor([local_sq], dtype=torch.float64)
                coll.all_reduce_(t, "sum")
                local_sq = float(t.item()) / n
            if (self._kf_step_count - 1) % self._kf_interval == 0 and n > 1:
                self.noise_scale_value = self._kf_cpu_ns.update(self._kf_bs, self._kf_bs * n, local_sq, big_sq)
                self._kf_gns_var.fill_(float(self.noise_scale_value))
                _publish_gauge(self.noise_scale_value)
            return loss
        loss = _SynchronousSGD.step(self, closure)
        sums = eng.read_sumsq(reset=True)  # [sum_r |g_r|^2, |g_avg|^2], device tensor
        if (self._kf_step_count - 1) % self._kf_interval == 0 and n > 1:
            st = torch.cuda.current_stream().cuda_stream
            eng.ext.gns_finalize(sums.data_ptr(), self._kf_gns_state.data_ptr(), self._kf_gns_out.data_ptr(), self._kf_bs,
                                 self._kf_bs * n, self._kf_alpha, float(n), st)
        return loss

    def noise_scale(self) -> float:
        """Latest noise-scale estimate (host value; syncs the stream on the GPU path)."""
        if self._kf_engine is not None:
            v = float(self._kf_gns_out[0].item())
            _publish_gauge(v)  # the device value reaches /metrics whenever somebody reads it
            return v
        return float(self.noise_scale_value)

    def noise_scale_tensor(self) -> Optional[torch.Tensor]:
        """Device tensor [noise scale, G_biased, S_biased] (no sync)."""
        return self._kf_gns_out if self._kf_engine is not None else None


def MonitorGradientNoiseScaleOptimizer(optimizer, named_parameters=None, device_batch_size: int = 1, monitor_interval: int = 1,
                                       alpha: float = 0.6, fused="auto", overlap: bool = True, bucket_mb=None, algo=None):
    """Reference signature: `(optimizer, device_batch_size, monitor_interval=1, name=None, use_locking=False)`."""
    return wrap_optimizer(optimizer, _MonitorGNS,
                          lambda o: o._kf_init_gns(named_parameters, device_batch_size, monitor_interval, alpha, fused, overlap, bucket_mb, algo))
