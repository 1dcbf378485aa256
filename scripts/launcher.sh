#!/bin/sh
# Shell helpers around kungfu-run: source this file, then `prun 4 python train.py`, `erun 2 python elastic.py`, `srun python x.py`
# (reference: scripts/launcher.sh — same three entry points and default flag sets).

_kf_root() { (cd "$(dirname "${KF_LAUNCHER_SH:-$0}")/.." && pwd); }

_kf_run() {
    if [ -n "$KUNGFU_RUN" ]; then echo "$KUNGFU_RUN"; return; fi
    if command -v kungfu-run >/dev/null 2>&1; then command -v kungfu-run; return; fi
    echo "$(_kf_root)/kungfu_b200/bin/kungfu-run"
}

kf_flags_default() { printf '%s\n' -q -logdir "${KF_LOGDIR:-logs}" -logfile kungfu-run.log; }

kf_flags_elastic() {
    port=${KF_CONFIG_PORT:-9100}
    kf_flags_default
    printf '%s\n' -w -builtin-config-port "$port" -config-server "http://127.0.0.1:$port/config"
}

kf_trace() { # run a command, print BEGIN/END with the wall time
    echo "BEGIN $*"
    _t0=$(date +%s)
    "$@"
    _rc=$?
    echo "END $*, took $(($(date +%s) - _t0))s (rc=$_rc)"
    echo
    return $_rc
}

prun() { # prun <np> <prog> [args]: static parallel run on this host
    _np=$1
    shift
    # shellcheck disable=SC2046
    kf_trace "$(_kf_run)" -np "$_np" $(kf_flags_default) "$@"
}

erun() { # erun <np> <prog> [args]: elastic run with the launcher's builtin config server
    _np=$1
    shift
    # shellcheck disable=SC2046
    kf_trace "$(_kf_run)" -np "$_np" -H "127.0.0.1:${KF_MAX_NP:-8}" $(kf_flags_elastic) "$@"
}

srun() { kf_trace "$@"; } # single-process mode (no launcher)
