"""Self-checking smoke programs: `kungfu-run -np 4 python -m kungfu_b200.examples` (reference:
`python -m kungfu.tensorflow.v1.examples` / `v2.examples`, srcs/python/kungfu/tensorflow/v1/examples/__main__.py:1-58)."""
