"""Launches the multi-GPU worker under torchrun on every available GPU count in {2, 4, 8}."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(n, what):
    port = 29600 + n
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp_gpu_worker.py")] + what
    env = dict(os.environ, KUNGFU_DEVICE_TIMEOUT_MS="15000")
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-6000:])
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"mp_{n}_{'_'.join(what)}.log"), "w") as f:
            f.write(r.stdout)
    except OSError:
        pass
    assert r.returncode == 0, r.stdout[-3000:]
    assert '"fails": 0' in r.stdout


def test_collectives_all_gpus():
    n = torch.cuda.device_count()
    _run(n if n in (2, 4, 8) else 2, ["allreduce", "bcast"])


def test_fused_optimizers_all_gpus():
    n = torch.cuda.device_count()
    _run(n if n in (2, 4, 8) else 2, ["ssgd"])


def test_sma_pair_all_gpus():
    n = torch.cuda.device_count()
    _run(n if n in (2, 4, 8) else 2, ["extra"])
