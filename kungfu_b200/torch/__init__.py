"""`kungfu.torch` parity (srcs/python/kungfu/torch/__init__.py) — and the Torch-first union of the reference's
TF-only features (SMA, PairAveraging, monitors, elastic control)."""
from ..python import (_get_cuda_index, current_cluster_size, current_local_rank, current_local_size, current_rank,
                      run_barrier)
from . import ops, optimizers

get_cuda_index = _get_cuda_index
broadcast_parameters = ops.broadcast_parameters


def nccl_built() -> bool:
    """The reference hard-returns False (its torch path never touches NCCL). Ours does not need NCCL at all:
    GPU collectives are our own kernels. Returns whether the NCCL *baseline* backend is usable."""
    try:
        import torch

        return bool(torch.cuda.is_available() and torch.distributed.is_nccl_available())
    except Exception:
        return False
