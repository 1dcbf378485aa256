#!/bin/sh
# GPU suite + smoke + headline bench on a B200 box (reference: scripts/tests/run-nccl-train-test.sh).
set -e
. "$(dirname "$0")/common.sh"
cd "$ROOT"
python -m pytest tests -x -q -m gpu "$@"
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 5
