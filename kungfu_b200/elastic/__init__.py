"""Elastic training helpers (reference: srcs/python/kungfu/tensorflow/hooks/elastic.py:14-87,
experimental/hook/elastic.py:11-116, policy/*.py, v1/datasets/adaptor.py)."""
from .datasets import BaseDatasetAdaptor, ClusterDatasetAdaptor, ElasticDataAdaptor, ExampleDatasetAdaptor
from .hooks import ElasticHook, ElasticState, KungFuElasticTrainHook, ResizeProfiler, ShrinkOnFailure, cuda_context_lost, is_peer_failure, sync_after_resize
from .policy import BasePolicy, PolicyHook, ScheduledBatchSizePolicy

__all__ = ["ElasticHook", "ElasticState", "KungFuElasticTrainHook", "ResizeProfiler", "sync_after_resize", "ShrinkOnFailure", "is_peer_failure", "cuda_context_lost", "BasePolicy",
           "PolicyHook", "ScheduledBatchSizePolicy", "ElasticDataAdaptor", "BaseDatasetAdaptor", "ExampleDatasetAdaptor",
           "ClusterDatasetAdaptor"]
