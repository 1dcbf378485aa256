"""All-reduce of ones (reference: examples/torch_simple_example.py).

    kungfu-run -np 4 python examples/torch_simple_example.py
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))  # run from a checkout without installing

import torch

import kungfu_b200.torch as kf

x = torch.ones(2, 3)
if torch.cuda.is_available() and kf.get_cuda_index() >= 0:
    torch.cuda.set_device(kf.get_cuda_index())
    x = x.cuda()
y = kf.ops.all_reduce_fn(x)
print(f"rank {kf.current_rank()}/{kf.current_cluster_size()}: {y.flatten().tolist()}")
