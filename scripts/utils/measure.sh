#!/bin/sh
# `measure <cmd …>`: run a command, report its wall time (reference: scripts/utils/measure.sh).
measure() {
    _m0=$(date +%s.%N)
    "$@"
    _mrc=$?
    _m1=$(date +%s.%N)
    echo "[measure] $* took $(echo "$_m1 - $_m0" | bc 2>/dev/null || python3 -c "print(round($_m1 - $_m0, 3))")s (rc=$_mrc)" >&2
    return $_mrc
}
