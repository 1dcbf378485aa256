#!/bin/sh
# Operator tests under the launcher for np 1..4 (reference: scripts/tests/run-op-tests.sh: test_operators.py / test_torch_ops.py per np).
set -e
. "$(dirname "$0")/common.sh"
cd "$ROOT"
for np in 1 2 3 4; do
    prun $np python tests/cpu_worker.py collectives
    prun $np python tests/cpu_worker.py p2p
    prun $np python tests/cpu_worker.py set_tree
    prun $np python tests/cpu_worker.py initializer
done
python -m pytest tests/test_runtime_unit.py tests/test_helpers.py -q "$@"
