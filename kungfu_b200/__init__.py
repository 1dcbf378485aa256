"""kungfu_b200 — a Blackwell-native adaptive distributed-training library for PyTorch.

Capabilities and API of KungFu (DingtongHan/KungFu-1): distributed optimizers (S-SGD, SMA, PairAveraging,
adaptive), online monitoring (gradient noise scale / variance, strategy stats), online control (barrier,
resize_cluster, set_tree), failure detection with auto-recovery and the `kungfu-run` launcher — re-designed for
one process per B200: a C++ control plane over sockets, peer-mapped HBM as the data plane, and hand-written
sm_100a kernels that fuse each collective with its adjacent compute over NVLink 5 / NVSwitch.

Sub-packages mirror the reference's Python layout (srcs/python/kungfu/): `python` (rank/size/barrier/resize),
`torch` (ops + optimizers), `cmd` (launcher entry + failure-detector signals), plus `ops`, `parallel`, `models`,
`utils`, `elastic`, `monitor`.
"""
__version__ = "0.1.0"


def build(verbose: bool = False, force: bool = False) -> None:
    """Build every native component in-tree (sm_100a kernels, CPU runtime, launcher)."""
    from . import _build

    _build.build_all(verbose=verbose, force=force)
