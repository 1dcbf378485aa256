"""Monitoring overhead: throughput with the gradient-noise-scale / variance monitors vs plain S-SGD.
Reference: benchmarks/monitoring/{README.md:1-17, benchmark.py:75-90}. Thin wrapper over the system benchmark."""
import os
import subprocess
import sys

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    script = os.path.join(os.path.dirname(here), "system", "benchmark_kungfu_torch.py")
    for opt in ("sync-sgd", "noise-scale", "variance"):
        subprocess.call([sys.executable, script, "--kf-optimizer", opt] + sys.argv[1:])
