"""Device plane: symmetric peer-mapped arenas (`mesh`), flat parameter / gradient layouts (`flat`) and the engine that launches
the fused optimizer kernels from autograd hooks (`engine`). Counterpart of the reference's NCCL controller / scheduler pair
(srcs/cpp/include/kungfu/nccl/{controller,scheduler,helper}.hpp); see DESIGN.md §2-3."""
from .flat import FlatLayout  # noqa: F401
from .mesh import GpuPlane, SymmetricArena  # noqa: F401
