#!/bin/sh
# Shared bits of the test drivers (reference: scripts/tests/common.sh).
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
export PYTHONPATH="$ROOT${PYTHONPATH:+:$PYTHONPATH}"
KUNGFU_RUN="$ROOT/kungfu_b200/bin/kungfu-run"
export KUNGFU_RUN
[ -x "$KUNGFU_RUN" ] || python -c "import kungfu_b200._build as b; b.build_all()"
KF_LAUNCHER_SH="$ROOT/scripts/launcher.sh"
. "$ROOT/scripts/launcher.sh"
export KF_LOGDIR=${KF_LOGDIR:-/tmp/kungfu-b200-logs}
