#!/bin/sh
# Reference: examples/mnist_elastic_eager/run.sh (kungfu-run -w, builtin config server, up to 4 workers on this host).
set -e
cd "$(dirname "$0")/../.."
. ./scripts/launcher.sh
KF_MAX_NP=${KF_MAX_NP:-4} erun 1 python examples/mnist_elastic_eager/main.py "$@"
