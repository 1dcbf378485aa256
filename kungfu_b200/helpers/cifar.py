"""CIFAR-10 / CIFAR-100 from the python-pickle distribution (`cifar-10-batches-py/`, `cifar-100-python/`).

Reference: helpers/cifar.py:24-103 (`Cifar10Loader`, `Cifar100Loader` with normalize / one_hot switches; images NHWC).
NHWC uint8 is also what the channels_last GPU pipeline wants: the batch crosses PCIe as bytes and is normalised on the device.
"""
from __future__ import annotations

import os
import pickle
from collections import namedtuple

import numpy as np

from .mnist import DataSet, to_onehot

_DEFAULT_DIR = os.path.join(os.path.expanduser("~"), "var", "data", "cifar")
Cifar10DataSets = namedtuple("Cifar10DataSets", "train test")
Cifar100DataSets = namedtuple("Cifar100DataSets", "train test")


def _unpickle(path: str) -> dict:
    with open(path, "rb") as f:
        return pickle.load(f, encoding="bytes")


class _Loader:
    classes = 10
    subdir = ""
    label_key = b"labels"
    train_files: tuple = ()
    test_file = ""
    bundle = Cifar10DataSets

    def __init__(self, data_dir: str = _DEFAULT_DIR, normalize: bool = False, one_hot: bool = False):
        self._dir, self._normalize, self._one_hot = data_dir, normalize, one_hot

    def _load(self, name: str) -> DataSet:
        rec = _unpickle(os.path.join(self._dir, self.subdir, name))
        data = np.asarray(rec[b"data"], np.uint8)
        images = data.reshape(data.shape[0], 3, 32, 32).transpose(0, 2, 3, 1)
        if self._normalize:
            images = (images / 255.0).astype(np.float32)
        labels = np.asarray(rec[self.label_key])
        if self._one_hot:
            labels = to_onehot(self.classes, labels)
        return DataSet(np.ascontiguousarray(images), labels)

    def load_train(self) -> DataSet:
        parts = [self._load(n) for n in self.train_files]
        return DataSet(np.concatenate([p.images for p in parts]), np.concatenate([p.labels for p in parts]))

    def load_test(self) -> DataSet:
        return self._load(self.test_file)

    def load_datasets(self):
        return self.bundle(self.load_train(), self.load_test())


class Cifar10Loader(_Loader):
    classes = 10
    subdir = "cifar-10-batches-py"
    train_files = tuple(f"data_batch_{i}" for i in range(1, 6))
    test_file = "test_batch"
    bundle = Cifar10DataSets


class Cifar100Loader(_Loader):
    classes = 100
    subdir = "cifar-100-python"
    label_key = b"fine_labels"
    train_files = ("train",)
    test_file = "test"
    bundle = Cifar100DataSets
