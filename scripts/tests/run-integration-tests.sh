#!/bin/sh
# Native integration programs under the launcher: np 1..4 x every strategy (reference: scripts/tests/run-integration-tests.sh:30-38 —
# fake-agent and test-p2p-apis over all strategies with a short time-out each).
set -e
. "$(dirname "$0")/common.sh"
cd "$ROOT"
BIN="$ROOT/kungfu_b200/bin"
for np in 1 2 3 4; do
    for strategy in STAR MULTI_STAR RING CLIQUE TREE BINARY_TREE BINARY_TREE_STAR MULTI_BINARY_TREE_STAR AUTO; do
        prun $np -strategy $strategy -timeout 60s "$BIN/kungfu-fake-agent" test
        prun $np -strategy $strategy -timeout 60s "$BIN/kungfu-test-public-apis"
    done
done
KUNGFU_CONFIG_ENABLE_MONITORING=true prun 4 -timeout 60s "$BIN/kungfu-test-public-apis"
"$BIN/kungfu-unit-tests"
