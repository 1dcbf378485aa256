#!/usr/bin/env bash
# Race / memory checking (the reference configures no sanitizers).
#   scripts/sanitize.sh thread      -> TSan build of the C++ runtime + launcher, then the C ABI integration program
#   scripts/sanitize.sh address     -> ASan/UBSan
#   scripts/sanitize.sh cuda        -> compute-sanitizer (memcheck, racecheck, synccheck) over the single-GPU kernel tests
set -euo pipefail
cd "$(dirname "$0")/.."
mode="${1:-thread}"
if [ "$mode" = "cuda" ]; then
    # one-rank forms of every collective / fused kernel (kungfu_b200/selftest.py), the optimizer, BN and transformer kernels
    sel="one_rank_mesh or allreduce_single_rank or flat_opt or cast or add_layernorm or bias_gelu or bn_"
    out="${2:-gpurun_out/compute_sanitizer}"
    mkdir -p "$(dirname "$out")"
    rc=0
    for tool in memcheck racecheck synccheck; do
        compute-sanitizer --tool "$tool" --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "$sel" > "$out.$tool.log" 2>&1 || rc=1
        # hazards inside PyTorch's own kernels (the fp32 reference side of a numerics test) are reported but are not ours
        ours=$(grep -E "Race reported|Invalid|Error:" "$out.$tool.log" | grep -c "kf::" || true)
        echo "compute-sanitizer $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$out.$tool.log" | tail -1) | reports naming kf:: kernels: $ours | $(grep -E 'passed|failed' "$out.$tool.log" | tail -1)"
    done
    exit $rc
fi
KUNGFU_SANITIZE="$mode" python -c "from kungfu_b200 import _build; _build.build_runtime(verbose=True)"
bin="kungfu_b200/bin-$mode"
R=$((32000 + RANDOM % 4000))
"$bin/kungfu-run" -q -np 4 -H 127.0.0.1:4 -port "$R" -port-range "$((R + 1))-$((R + 50))" -strategy RING "$bin/kungfu-fake-agent" test
"$bin/kungfu-run" -q -np 3 -H 127.0.0.1:3 -port "$R" -port-range "$((R + 1))-$((R + 50))" -strategy BINARY_TREE "$bin/kungfu-fake-agent" bench slp
# the watch-mode launcher paths of round 2 (loss announcements, hang watch, refill from the standby pool) with plain Python workers: the
# instrumented process is the launcher
export PYTHONPATH="$PWD" KUNGFU_CONFIG_LOG_LEVEL=WARN
KUNGFU_TEST_HANG=1 KUNGFU_MONITOR_PORT=$((17000 + RANDOM % 2000)) "$bin/kungfu-run" -q -w -shrink-on-failure -auto-recover 3s -np 3 -H 127.0.0.1:4 \
    -port "$R" -port-range "$((R + 1))-$((R + 50))" -builtin-config-port $((9200 + RANDOM % 90)) python tests/elastic_opt_worker.py shrink 1@3 8
KUNGFU_TEST_REFILL=1 "$bin/kungfu-run" -q -w -shrink-on-failure -refill -standby 1 -np 3 -H 127.0.0.1:4 \
    -port "$R" -port-range "$((R + 1))-$((R + 50))" -builtin-config-port $((9200 + RANDOM % 90)) python tests/elastic_opt_worker.py shrink 2@4 40
# ... and a fully instrumented job (launcher AND workers): the C-ABI agent loses a worker and recovers
"$bin/kungfu-run" -q -w -shrink-on-failure -np 3 -H 127.0.0.1:4 -port "$R" -port-range "$((R + 1))-$((R + 50))" \
    -builtin-config-port $((9200 + RANDOM % 90)) "$bin/kungfu-fake-agent" shrink 1@5
# fully instrumented native tools: the public C++ API end to end (monitoring endpoint included), concurrent p2p pulls, an elastic job that
# grows and shrinks 1 -> 2 -> 4 -> 2 -> 1 (resize protocol, runners starting / reaping workers), four peers as threads of one process
KUNGFU_CONFIG_ENABLE_MONITORING=true "$bin/kungfu-run" -q -np 4 -H 127.0.0.1:4 -port "$R" -port-range "$((R + 1))-$((R + 50))" "$bin/kungfu-test-public-apis"
"$bin/kungfu-run" -q -np 3 -H 127.0.0.1:3 -port "$R" -port-range "$((R + 1))-$((R + 50))" "$bin/kungfu-bench-p2p" -model slp -mode par -epochs 5
"$bin/kungfu-run" -q -w -np 1 -H 127.0.0.1:4 -port "$R" -port-range "$((R + 1))-$((R + 50))" -builtin-config-port $((9200 + RANDOM % 90)) \
    "$bin/kungfu-fake-adaptive-trainer" -max-step 14 -schedule 2:2,4:4,7:2,10:1 -slow-ms 20 -fast-ms 5
"$bin/kungfu-fake-inproc-trainer" -np 4 -model slp -steps 20
# every tensor of a model in flight at once over 32 threads; workers that die in the loader are restarted by the (instrumented) launcher;
# the outside tools against a live job: ping a worker, resize through the config server, stop the last worker
"$bin/kungfu-run" -q -np 3 -H 127.0.0.1:3 -port "$R" -port-range "$((R + 1))-$((R + 50))" -strategy BINARY_TREE_STAR "$bin/kungfu-bench-allreduce" -model resnet50 -mode par -epochs 2 -warmup 1
"$bin/kungfu-run" -q -np 3 -H 127.0.0.1:3 -port "$R" -port-range "$((R + 1))-$((R + 50))" "$bin/kungfu-bench-allreduce" -model slp -rand-failure 0.5
C=$((9200 + RANDOM % 90))
"$bin/kungfu-run" -q -w -np 1 -H 127.0.0.1:3 -port "$R" -port-range "$((R + 1))-$((R + 50))" -builtin-config-port "$C" \
    "$bin/kungfu-fake-adaptive-trainer" -max-step 100000 -slow-ms 20 -fast-ms 5 &
job=$!
"$bin/kungfu-test-util" -wait "127.0.0.1:$((R + 1))" -timeout 20s
"$bin/kungfu-test-util" -ping "127.0.0.1:$((R + 1))" -n 3
"$bin/kungfu-config-client" -url "http://127.0.0.1:$C/config" -size 3
sleep 1
"$bin/kungfu-config-client" -url "http://127.0.0.1:$C/config" -random -min 1 -max 3 -period 300ms -duration 2s -final 1
sleep 1
"$bin/kungfu-test-util" -kill "127.0.0.1:$((R + 1))"
wait "$job"
# the C++ unit tests under the same instrumentation: in-process multi-peer worlds (threads), the epoch fence, a world that loses a peer
"$bin/kungfu-unit-tests"
echo "sanitizer run ($mode) passed"
