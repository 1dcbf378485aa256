"""MNIST from the four idx files (`train|t10k-{images-idx3,labels-idx1}-ubyte[.gz]`).

Reference: helpers/mnist.py:18-48 (`load_mnist_data(data_dir, prefix, normalize, one_hot, padded)`, `load_datasets`).
`synthetic` builds a deterministic stand-in with the same shapes for boxes without the files (the CI of this repo has no
network); it is a linearly separable problem, so convergence tests remain meaningful.
"""
from __future__ import annotations

import os
from collections import namedtuple

import numpy as np

from .idx import read_idx_file

DataSet = namedtuple("DataSet", "images labels")
MnistDataSets = namedtuple("MnistDataSets", "train test")
_PREFIXES = ("train", "t10k")


def to_onehot(k: int, a: np.ndarray) -> np.ndarray:
    out = np.zeros((a.shape[0], k), np.float32)
    out[np.arange(a.shape[0]), a.astype(np.int64)] = 1.0
    return out


def _find(data_dir: str, stem: str) -> str:
    for cand in (stem, stem + ".gz", stem.replace("-idx", ".idx"), stem.replace("-idx", ".idx") + ".gz"):
        p = os.path.join(data_dir, cand)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(os.path.join(data_dir, stem))


def _finish(images, labels, normalize, one_hot, padded):
    if padded:  # LeNet-style 32x32x1
        images = np.pad(images.reshape(images.shape[0], 28, 28, 1), ((0, 0), (2, 2), (2, 2), (0, 0)), "constant")
    if normalize:
        images = (images / 255.0).astype(np.float32)
    if one_hot:
        labels = to_onehot(10, labels)
    return DataSet(images, labels)


def load_mnist_data(data_dir: str, prefix: str, normalize: bool, one_hot: bool, padded: bool = False) -> DataSet:
    if prefix not in _PREFIXES:
        raise ValueError("prefix must be %s" % " | ".join(_PREFIXES))
    images = read_idx_file(_find(data_dir, prefix + "-images-idx3-ubyte"))
    labels = read_idx_file(_find(data_dir, prefix + "-labels-idx1-ubyte"))
    return _finish(images, labels, normalize, one_hot, padded)


def load_datasets(data_dir: str, normalize: bool = False, one_hot: bool = False, padded: bool = False) -> MnistDataSets:
    return MnistDataSets(load_mnist_data(data_dir, "train", normalize, one_hot, padded),
                         load_mnist_data(data_dir, "t10k", normalize, one_hot, padded))


def synthetic(n_train: int = 6000, n_test: int = 1000, seed: int = 0, normalize: bool = False, one_hot: bool = False,
              padded: bool = False) -> MnistDataSets:
    """Ten fixed 28x28 uint8 prototypes plus noise; same dtypes and shapes as the real files."""
    rng = np.random.RandomState(seed)
    protos = rng.randint(0, 256, size=(10, 28, 28)).astype(np.float32)

    def make(n):
        labels = rng.randint(0, 10, size=n).astype(np.uint8)
        images = np.clip(0.6 * protos[labels] + 0.4 * rng.randint(0, 256, size=(n, 28, 28)), 0, 255).astype(np.uint8)
        return _finish(images, labels, normalize, one_hot, padded)

    return MnistDataSets(make(n_train), make(n_test))


def write_datasets(data_dir: str, sets: MnistDataSets) -> None:
    """Store a (synthetic or real) set in the canonical idx layout so that `load_datasets` reads it back."""
    from .idx import write_idx_file

    os.makedirs(data_dir, exist_ok=True)
    for prefix, ds in zip(_PREFIXES, sets):
        write_idx_file(os.path.join(data_dir, prefix + "-images-idx3-ubyte"), np.asarray(ds.images, np.uint8).reshape(-1, 28, 28))
        write_idx_file(os.path.join(data_dir, prefix + "-labels-idx1-ubyte"), np.asarray(ds.labels, np.uint8))
