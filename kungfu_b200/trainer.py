"""High-level training loop helper: the call a user makes for one end-to-end step.

`ImageTrainer.step_async(images_u8, labels)` takes a *pinned host* uint8 NHWC batch, copies it to the device on a copy
stream (double-buffered, overlapping the previous step's compute), normalises+casts it with our fused kernel, replays
the captured forward+backward+fused-optimizer CUDA graph and copies the loss back to pinned host memory.

The reference has no such helper — its benchmark keeps one synthetic batch resident on the device
(benchmarks/system/benchmark_kungfu_torch.py:91-103); this is the path our `e2e` number is measured through.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn.functional as F

from .runtime import get_runtime
from .utils import dtypes

IMAGENET_MEAN = (0.485 * 255, 0.456 * 255, 0.406 * 255)
IMAGENET_STD = (0.229 * 255, 0.224 * 255, 0.225 * 255)


class LossFuture:
    def __init__(self, host: torch.Tensor, event: Optional[torch.cuda.Event]) -> None:
        self._host = host
        self._event = event

    def result(self) -> float:
        if self._event is None:  # host execution: the value is already there
            return float(self._host.item())
        self._event.synchronize()
        return float(self._host.item())


class ImageTrainer:
    def __init__(self, model: torch.nn.Module, optimizer, batch_size: int, image_size: int = 224,
                 dtype: torch.dtype = torch.bfloat16, loss_fn: Optional[Callable] = None, use_graph: bool = True,
                 warmup_iters: int = 3, device: Optional[torch.device] = None, capture_error_mode: str = "global") -> None:
        self.capture_error_mode = capture_error_mode
        rt = get_runtime()
        self.rt = rt
        self.device = device or (rt.gpu.device if rt.has_gpu() else torch.device("cpu"))
        self.model = model
        self.optimizer = optimizer
        self.batch = batch_size
        self.dtype = dtype
        self.loss_fn = loss_fn or (lambda out, y: F.cross_entropy(out.float(), y))
        self.use_graph = use_graph and self.device.type == "cuda"
        H = W = image_size
        dev = self.device
        # static device inputs in channels_last (NHWC bytes == the uint8 wire format)
        self.x = torch.zeros((batch_size, 3, H, W), dtype=dtype, device=dev).contiguous(memory_format=torch.channels_last)
        self.y = torch.zeros((batch_size,), dtype=torch.int64, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_error: Optional[str] = None
        self.launches_per_step = 0  # launches of OUR kernels inside one step
        if dev.type == "cuda":
            self.ext = rt.gpu.ext
            self.copy_stream = torch.cuda.Stream(device=dev)
            self.u8 = [torch.empty((batch_size, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
            self.yb = [torch.empty((batch_size,), dtype=torch.int64, device=dev) for _ in range(2)]
            self.u8_free = [torch.cuda.Event(), torch.cuda.Event()]
            for e in self.u8_free:
                e.record()
            self.mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32, device=dev)
            self.invstd = torch.tensor([1.0 / s for s in IMAGENET_STD], dtype=torch.float32, device=dev)
            self.loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
            self._i = 0
            self._prepare(warmup_iters)

    # ---------------------------------------------------------------------------------------------
    def _train_body(self) -> None:
        self.optimizer.zero_grad()
        out = self.model(self.x)
        loss = self.loss_fn(out, self.y)
        loss.backward()
        self.optimizer.step()
        self.loss.copy_(loss.detach())

    def _prepare(self, warmup_iters: int) -> None:
        """Eager warm-up (cuDNN autotune, lazy state) then capture the whole step into one CUDA graph."""
        torch.cuda.synchronize()
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup_iters)):
                c0 = self.ext.launch_count()
                self._train_body()
                self.launches_per_step = self.ext.launch_count() - c0
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if not self.use_graph:
            return
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s, capture_error_mode=self.capture_error_mode):
                self._train_body()
            torch.cuda.synchronize()
            self.graph = g
        except Exception as e:  # capture is an optimisation; fall back to eager and say so
            self.graph = None
            self.graph_error = f"{type(e).__name__}: {e}"
            torch.cuda.synchronize()

    # ---------------------------------------------------------------------------------------------
    def device_step(self) -> torch.Tensor:
        """One training step on the resident device batch (kernel-only path)."""
        if self.graph is not None:
            self.graph.replay()
            eng = getattr(self.optimizer, "engine", None)
            if eng is not None and eng.nranks > 1 and eng.arena.mesh.status()[0] != 0:
                eng.check()  # replays bypass the optimizer's python: poll the host-mapped status word of the device-side waits here
        else:
            self._train_body()
        return self.loss

    def load_batch(self, images_u8: torch.Tensor, labels: torch.Tensor) -> None:
        """H2D (copy stream) + fused normalise/cast into the static input (main stream)."""
        i = self._i & 1
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.u8_free[i])
            self.u8[i].copy_(images_u8, non_blocking=True)
            self.yb[i].copy_(labels, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        cur.wait_event(ready)
        npix = self.batch * self.x.shape[2] * self.x.shape[3]
        self.ext.u8_normalize(self.u8[i].data_ptr(), self.x.data_ptr(), dtypes.dtype_code(self.dtype), npix,
                              self.mean.data_ptr(), self.invstd.data_ptr(), cur.cuda_stream)
        self.y.copy_(self.yb[i], non_blocking=True)
        self.u8_free[i].record(cur)

    def step_async(self, images_u8: torch.Tensor, labels: torch.Tensor) -> LossFuture:
        """Full end-to-end step from pinned host memory; returns a future for the loss."""
        if self.device.type != "cuda":
            return self._step_host(images_u8, labels)
        self.load_batch(images_u8, labels)
        self.device_step()
        i = self._i & 1
        self.loss_host[i].copy_(self.loss, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._i += 1
        return LossFuture(self.loss_host[i], ev)

    def _step_host(self, images_u8: torch.Tensor, labels: torch.Tensor) -> LossFuture:
        """The same step without a GPU (CI, debugging): normalise with torch ops, run the model and the optimizer eagerly."""
        mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32)
        std = torch.tensor(IMAGENET_STD, dtype=torch.float32)
        x = (images_u8.to(torch.float32) - mean) / std  # NHWC
        self.x.copy_(x.permute(0, 3, 1, 2).to(self.dtype))
        self.y.copy_(labels)
        self._train_body()
        return LossFuture(self.loss.detach().clone(), None)

    def step(self, images_u8: torch.Tensor, labels: torch.Tensor) -> float:
        return self.step_async(images_u8, labels).result()

    @property
    def h2d_bytes_per_step(self) -> int:
        if self.device.type != "cuda":
            return 0
        return self.u8[0].numel() + self.yb[0].numel() * 8

    @property
    def d2h_bytes_per_step(self) -> int:
        return 4


class GraphedStep:
    """Capture one whole training step (zero_grad -> forward -> backward -> optimizer.step) of ANY model into a CUDA graph.

        step = GraphedStep(lambda: train_body(), optimizer)   # eager warm-up, then capture
        step()                                                 # replay (or eager when capture is disabled / failed)

    `body` must use static input tensors (refill them in place between calls). What changes from step to step lives on the
    device — learning rate / step counters (engine hyper block), the PairAveraging target (a device word) — so one graph serves
    the whole run; `optimizer.prepare_replay()` (if present) refreshes those words before every replay.
    """

    def __init__(self, body: Callable[[], None], optimizer=None, warmup_iters: int = 3, use_graph: bool = True,
                 capture_error_mode: str = "global") -> None:
        self.body = body
        self.optimizer = optimizer
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_error: Optional[str] = None
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup_iters)):
                body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if not use_graph:
            return
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s, capture_error_mode=capture_error_mode):
                body()
            torch.cuda.synchronize()
            self.graph = g
        except Exception as e:  # capture is an optimisation; fall back to eager and say so
            self.graph_error = f"{type(e).__name__}: {e}"
            torch.cuda.synchronize()

    def __call__(self) -> None:
        if self.graph is None:
            self.body()
            return
        prep = getattr(self.optimizer, "prepare_replay", None)
        if prep is not None:
            prep()
        self.graph.replay()
        eng = getattr(self.optimizer, "engine", None)
        if eng is not None and eng.nranks > 1 and eng.arena.mesh.status()[0] != 0:
            eng.check()
