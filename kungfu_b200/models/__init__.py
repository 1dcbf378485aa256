"""Model zoo used by the benchmarks and tests: ResNet (headline), BERT (PairAveraging), SLP (plumbing)."""
from . import bert, resnet, slp  # noqa: F401
