"""Single-layer perceptron (784 -> 10 softmax): the model of the reference's plumbing / convergence-equivalence tests
(srcs/python/kungfu/tensorflow/v1/benchmarks/mnist.py:6-9, tests/python/integration/test_mnist_slp.py, examples/
torch_mnist_example.py) and of BASELINE config #1 ("SLP MNIST SynchronousSGDOptimizer world_size=2 on CPU")."""
from __future__ import annotations

import torch
import torch.nn as nn


class SLP(nn.Module):
    def __init__(self, inputs: int = 784, classes: int = 10, zero_init: bool = True) -> None:
        super().__init__()
        self.fc = nn.Linear(inputs, classes)
        if zero_init:  # the reference test zero-initialises so that results are exactly reproducible
            nn.init.zeros_(self.fc.weight)
            nn.init.zeros_(self.fc.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc(x.flatten(1))


def synthetic_mnist(n: int = 6000, seed: int = 0):
    """MNIST-shaped synthetic data (there is no dataset access here): 10 noisy class prototypes, linearly separable."""
    g = torch.Generator().manual_seed(seed)
    protos = torch.rand(10, 28 * 28, generator=g)
    y = torch.randint(0, 10, (n,), generator=g)
    x = (protos[y] + 0.35 * torch.randn(n, 28 * 28, generator=g)).clamp(0, 1).view(n, 1, 28, 28)
    return x, y
