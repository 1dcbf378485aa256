"""Distributed optimizers (reference: srcs/python/kungfu/tensorflow/optimizers/__init__.py and
srcs/python/kungfu/torch/optimizers/__init__.py)."""
from .ada_sgd import AdaptiveSGDOptimizer, AdaSGDHook
from .async_sgd import PairAveragingOptimizer, get_random_peer
from .grad_noise_scale import MonitorGradientNoiseScaleOptimizer
from .grad_variance import MonitorGradientVarianceOptimizer
from .sma_sgd import SynchronousAveragingOptimizer
from .sync_sgd import SynchronousSGDOptimizer

__all__ = ["SynchronousSGDOptimizer", "SynchronousAveragingOptimizer", "PairAveragingOptimizer", "AdaptiveSGDOptimizer",
           "MonitorGradientNoiseScaleOptimizer", "MonitorGradientVarianceOptimizer", "AdaSGDHook", "get_random_peer"]
