"""Distributed optimizers (reference: srcs/python/kungfu/tensorflow/optimizers/__init__.py and
srcs/python/kungfu/torch/optimizers/__init__.py)."""
from .async_sgd import PairAveragingOptimizer
from .sma_sgd import SynchronousAveragingOptimizer
from .sync_sgd import SynchronousSGDOptimizer

__all__ = ["SynchronousSGDOptimizer", "SynchronousAveragingOptimizer", "PairAveragingOptimizer"]
