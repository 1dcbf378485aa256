"""In-process launch of a KungFu cluster without kungfu-run (reference: examples/multiprocess/torch_multiprocess.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "../..")))  # run from a checkout without installing


def worker(rank):
    import kungfu_b200.torch as kf

    y = kf.ops.all_reduce_fn(torch.ones(4) * (rank + 1))
    print(f"rank {rank}/{kf.current_cluster_size()}: {y.tolist()}")


if __name__ == "__main__":
    from kungfu_b200.cmd import launch_multiprocess

    launch_multiprocess(worker, 4)
