"""Interference detection -> strategy switch (reference: experimental/adapt_strategy/adapt_strategy.py:176-203).

Every window: monitored all-reduces feed the throughput statistics; when a majority of peers see throughput below 80 % of
the reference window, switch to an alternative tree (here: the minimum spanning tree over measured peer latencies).

    kungfu-run -np 4 -strategy STAR python examples/adapt_strategy.py
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))  # run from a checkout without installing

import torch

import kungfu_b200.python as kfp
import kungfu_b200.torch as kf
from kungfu_b200 import ops


def main():
    n = kf.current_cluster_size()
    star = [0] * n
    ops.set_tree(star)
    x = torch.ones(1 << 20)
    for window in range(6):
        for _ in range(10):
            ops.monitored_all_reduce(x, tree=star)
        kfp.calc_stats()
        kfp.log_stats()
        if kfp.check_interference():
            father = ops.minimum_spanning_tree_father(ops.get_peer_latencies())
            ops.set_tree(father)
            star = father
            if kf.current_rank() == 0:
                print(f"window {window}: interference detected, switched to tree {father}")
    kfp.print_strategy_stats()


if __name__ == "__main__":
    main()
