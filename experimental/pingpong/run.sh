#!/bin/sh
# Single-worker and 4-worker runs of the ping-pong policy-gradient toy (reference: experimental/atari_pingpong/run.sh).
set -e
cd "$(dirname "$0")"
ROOT=$(cd ../.. && pwd)
KUNGFU_RUN=$ROOT/kungfu_b200/bin/kungfu-run
[ -x "$KUNGFU_RUN" ] || python -c "import sys; sys.path.insert(0, '$ROOT'); import kungfu_b200._build as b; b.build_all()"
CKPT=${CKPT:-/tmp/pingpong.npz}
TOTAL_BATCH=${TOTAL_BATCH:-8}
NP=${NP:-4}

single_train() { PYTHONUNBUFFERED=1 python ./pingpong.py --batch-size "$TOTAL_BATCH" "$@"; }

parallel_train() {
    python ./pingpong.py --init --checkpoint "$CKPT"
    "$KUNGFU_RUN" -timeout 10m -np "$NP" python ./pingpong.py --checkpoint "$CKPT" --batch-size $((TOTAL_BATCH / NP)) "$@"
}

case "${1:-parallel}" in
single) shift; single_train "$@" ;;
*) [ $# -gt 0 ] && shift; parallel_train "$@" ;;
esac
