#!/bin/sh
# Elastic schedules, natively and through the Python API (reference: scripts/tests/run-tensorflow-resize-test.sh, run-elastic-test.sh).
set -e
. "$(dirname "$0")/common.sh"
cd "$ROOT"
BIN="$ROOT/kungfu_b200/bin"
KF_MAX_NP=8 erun 1 "$BIN/kungfu-fake-adaptive-trainer" -max-step 24 -schedule 2:2,5:4,8:8,11:4,14:2,17:1
KF_MAX_NP=4 erun 1 python examples/torch_elastic/torch_mnist_elastic.py
KF_MAX_NP=4 erun 1 python examples/mnist_elastic_policy/main.py --schedule 10:2,20:3,40:4,50:1
KF_MAX_NP=4 erun 1 python benchmarks/system/benchmark_kungfu_elastic.py --no-cuda --model slp --num-iters 20 --schedule 5:2,10:4,15:1
