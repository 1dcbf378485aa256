#!/bin/sh
# CPU suite, as the CI runs it (reference: scripts/tests/run-python-tests.sh + run-op-tests.sh + run-optimizer-tests.sh).
set -e
. "$(dirname "$0")/common.sh"
cd "$ROOT"
python -m pytest tests -x -q -m "not gpu" "$@"
