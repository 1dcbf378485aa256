"""ImageNet-style input pipeline.

Reference: helpers/imagenet.py:11-164 parses TFRecords and does the Inception-style training augmentation (random crop with
aspect/area ranges, flip, colour distortion) in the TF graph. There is no TF here; the B200 equivalent keeps the host work to
"read bytes" and does the augmentation ON THE GPU on a whole uint8 batch at once (one crop/flip per sample through a batched
affine grid, brightness/contrast/saturation jitter as fused elementwise ops), after the H2D copy that the trainer's copy
stream already overlaps with compute.

`create_dataset(data_dir, batch_size, n)` reads pre-decoded shards `train-XXXXX-of-YYYYY.npz` (arrays `images` uint8 NHWC and
`labels`) in the layout written by `write_shard`; `n` limits the number of shards like the reference's argument.
"""
from __future__ import annotations

import glob
import os
from typing import Iterator, List, Tuple

import numpy as np
import torch

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def write_shard(path: str, images: np.ndarray, labels: np.ndarray) -> None:
    assert images.dtype == np.uint8 and images.ndim == 4 and images.shape[-1] == 3, "images must be uint8 NHWC"
    np.savez(path, images=images, labels=np.asarray(labels, np.int64))


def list_shards(data_dir: str, subset: str = "train", n: int | None = None) -> List[str]:
    names = sorted(glob.glob(os.path.join(data_dir, f"{subset}-*-of-*.npz")))
    return names if n is None else names[:n]


def create_dataset_from_files(filenames: List[str], batch_size: int, drop_remainder: bool = True,
                              shard: Tuple[int, int] = (0, 1)) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
    """Yield pinned (uint8 NHWC images, int64 labels) batches; `shard=(rank, size)` strides samples across workers."""
    rank, size = shard
    carry_x: list = []
    carry_y: list = []
    have = 0
    for name in filenames:
        with np.load(name) as z:
            x, y = z["images"][rank::size], z["labels"][rank::size]
        carry_x.append(x)
        carry_y.append(y)
        have += len(x)
        while have >= batch_size:
            X, Y = np.concatenate(carry_x), np.concatenate(carry_y)
            bx, by = X[:batch_size], Y[:batch_size]
            carry_x, carry_y, have = [X[batch_size:]], [Y[batch_size:]], len(X) - batch_size
            yield _pin(torch.from_numpy(np.ascontiguousarray(bx))), _pin(torch.from_numpy(np.ascontiguousarray(by)))
    if have and not drop_remainder:
        X, Y = np.concatenate(carry_x), np.concatenate(carry_y)
        yield _pin(torch.from_numpy(np.ascontiguousarray(X))), _pin(torch.from_numpy(np.ascontiguousarray(Y)))


def create_dataset(data_dir: str, batch_size: int = 32, n: int = 1, subset: str = "train", shard: Tuple[int, int] = (0, 1)):
    return create_dataset_from_files(list_shards(data_dir, subset, n), batch_size, shard=shard)


def _pin(t: torch.Tensor) -> torch.Tensor:
    return t.pin_memory() if torch.cuda.is_available() else t


@torch.no_grad()
def train_augment(batch_u8: torch.Tensor, out_hw: Tuple[int, int] = (224, 224), area=(0.05, 1.0), aspect=(0.75, 1.33),
                  color: bool = True, generator: torch.Generator | None = None, dtype=torch.bfloat16) -> torch.Tensor:
    """Random-resized-crop + horizontal flip + colour jitter + normalisation of a uint8 NHWC batch, entirely on its device.
    Returns NCHW-shaped, channels_last-strided `dtype` (what the ResNet stem consumes)."""
    B, H, W, C = batch_u8.shape
    dev = batch_u8.device
    u = lambda *s: torch.rand(*s, device=dev, generator=generator)  # noqa: E731
    frac = area[0] + (area[1] - area[0]) * u(B)
    logr = torch.log(torch.tensor(aspect, device=dev))
    ratio = torch.exp(logr[0] + (logr[1] - logr[0]) * u(B))
    sw = torch.sqrt(frac * ratio).clamp(max=1.0)  # crop width / height as a fraction of the image
    sh = torch.sqrt(frac / ratio).clamp(max=1.0)
    cx = (u(B) * 2 - 1) * (1 - sw)
    cy = (u(B) * 2 - 1) * (1 - sh)
    flip = torch.where(u(B) < 0.5, -1.0, 1.0)
    theta = torch.zeros(B, 2, 3, device=dev)
    theta[:, 0, 0], theta[:, 0, 2], theta[:, 1, 1], theta[:, 1, 2] = sw * flip, cx, sh, cy
    x = batch_u8.permute(0, 3, 1, 2).float().div_(255.0)  # NCHW view over NHWC memory = channels_last
    grid = torch.nn.functional.affine_grid(theta, (B, C, out_hw[0], out_hw[1]), align_corners=False)
    x = torch.nn.functional.grid_sample(x, grid, mode="bilinear", padding_mode="border", align_corners=False)
    if color:
        bright = (u(B, 1, 1, 1) - 0.5) * (64.0 / 255.0)
        contrast = 0.5 + u(B, 1, 1, 1)
        sat = 0.5 + u(B, 1, 1, 1)
        gray = (0.299 * x[:, 0:1] + 0.587 * x[:, 1:2] + 0.114 * x[:, 2:3])
        x = gray + (x - gray) * sat
        mean = x.mean(dim=(1, 2, 3), keepdim=True)
        x = ((x - mean) * contrast + mean + bright).clamp_(0.0, 1.0)
    mean = torch.tensor(MEAN, device=dev).view(1, 3, 1, 1)
    std = torch.tensor(STD, device=dev).view(1, 3, 1, 1)
    return ((x - mean) / std).to(dtype).contiguous(memory_format=torch.channels_last)


@torch.no_grad()
def eval_preprocess(batch_u8: torch.Tensor, out_hw: Tuple[int, int] = (224, 224), central_fraction: float = 0.875,
                    dtype=torch.bfloat16) -> torch.Tensor:
    """Central crop + resize + normalisation (the reference's eval branch)."""
    B, H, W, C = batch_u8.shape
    ch, cw = int(round(H * central_fraction)), int(round(W * central_fraction))
    top, left = (H - ch) // 2, (W - cw) // 2
    x = batch_u8[:, top:top + ch, left:left + cw].permute(0, 3, 1, 2).float().div_(255.0)
    x = torch.nn.functional.interpolate(x, size=out_hw, mode="bilinear", align_corners=False)
    mean = torch.tensor(MEAN, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(STD, device=x.device).view(1, 3, 1, 1)
    return ((x - mean) / std).to(dtype).contiguous(memory_format=torch.channels_last)


def record_to_labeled_image(record: dict, training: bool = True) -> Tuple[torch.Tensor, int]:
    """One decoded record ({'image': uint8 HWC array, 'label': int}) -> (augmented CHW tensor, label)."""
    img = torch.as_tensor(record["image"]).unsqueeze(0)
    out = train_augment(img, dtype=torch.float32) if training else eval_preprocess(img, dtype=torch.float32)
    return out[0], int(record["label"])
