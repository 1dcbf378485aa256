#!/bin/sh
# Elastic schedules under the watch-mode launcher (reference: scripts/tests/run-elastic-test.sh, run-tensorflow-resize-test.sh).
set -e
. "$(dirname "$0")/common.sh"
cd "$ROOT"
KF_MAX_NP=4 erun 1 python examples/torch_elastic/torch_mnist_elastic.py "$@"
