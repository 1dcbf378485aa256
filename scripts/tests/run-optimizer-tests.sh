#!/bin/sh
# The optimizers for np 1..4, values checked against closed forms (reference: scripts/tests/run-optimizer-tests.sh runs them unchecked).
set -e
. "$(dirname "$0")/common.sh"
cd "$ROOT"
python -m pytest tests -q -m "not gpu" -k "optimizer or ssgd or sma or pair or noise or variance or ada" "$@"
