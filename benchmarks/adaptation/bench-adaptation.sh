#!/bin/sh
# Elastic adaptation benchmark on one box (reference: benchmarks/adaptation/bench-adaptation.sh: 2 -> ... -> 16 -> 1 on 16 slots).
set -e
cd "$(dirname "$0")/../.."
CAP=${CAP:-8}
KUNGFU_RUN=${KUNGFU_RUN:-kungfu_b200/bin/kungfu-run}
run() {
    init_np=$1
    shift
    "$KUNGFU_RUN" -q -H 127.0.0.1:"$CAP" -np "$init_np" -timeout 4m -port-range 20001-20064 -w -builtin-config-port "${CFG_PORT:-9177}" "$@"
}
run 2 python benchmarks/adaptation/elastic_resize.py --schedule "2:5,$CAP:5,1:5,$CAP:5,1:5" --batch-size "${BATCH:-8}" "$@"
