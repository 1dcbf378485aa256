#!/bin/sh
# Convergence-equivalence: the SLP result must not depend on the number of workers (reference: scripts/tests/
# run-train-tests.sh:76-81 — np in {1,2,3,4} for each global batch size).
set -e
. "$(dirname "$0")/common.sh"
cd "$ROOT"
for bs in 48 120; do
    for np in 1 2 3 4; do
        prun $np python examples/torch_mnist_example.py --epochs 1 --batch-size $bs
    done
done
