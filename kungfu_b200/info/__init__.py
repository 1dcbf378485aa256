"""`python -m kungfu_b200.info` (or `python -m kungfu.info`): what this installation was built with and what it can see.

Reference: srcs/python/kungfu/info/__main__.py:1-26 prints the CUDA, NCCL and TensorFlow versions. Here: framework version,
torch / CUDA build, the native pieces that are present, the GPUs and whether NVLS multicast is supported on them.
"""
from __future__ import annotations

import json

from ..utils import info as _info


def collect() -> dict:
    d = dict(_info())
    try:
        import torch

        if torch.cuda.is_available():
            d["gpus"] = [torch.cuda.get_device_name(i) for i in range(torch.cuda.device_count())]
            d["compute_capability"] = ".".join(map(str, torch.cuda.get_device_capability(0)))
            try:
                d["nccl"] = ".".join(map(str, torch.cuda.nccl.version()))
            except Exception:
                pass
    except Exception as e:  # pragma: no cover
        d["gpu_probe_error"] = repr(e)
    return d


def main(argv=None) -> int:
    d = collect()
    print("kungfu_b200 %s" % d.get("version"))
    print("torch %s (CUDA build %s), CUDA available: %s" % (d.get("torch"), d.get("cuda_build"), d.get("cuda_available")))
    print("native runtime (_kfrt): %s, CUDA kernels (_kfcuda): %s" % (d.get("native_runtime"), d.get("cuda_kernels")))
    for i, g in enumerate(d.get("gpus", [])):
        print("GPU %d: %s" % (i, g))
    print(json.dumps(d))
    return 0
