"""Communication micro-benchmarks (reference: `python -m kungfu.tensorflow.v1.benchmarks`,
srcs/python/kungfu/tensorflow/v1/benchmarks/__main__.py)."""
