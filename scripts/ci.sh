#!/usr/bin/env bash
# What CI runs on a machine without a GPU (see .github/workflows/ci.yaml): build every native component (the CUDA kernels are
# cross-compiled for sm_100a), the C++ unit tests, the CPU pytest suite (multi-process integration under kungfu-run included) and
# the ThreadSanitizer pass over the runtime. Reference counterparts: `./configure --build-tests && make && make test`, `go test
# ./...`, scripts/tests/run-*.sh (ci.yaml:29-64).
set -euo pipefail
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
python scripts/check_names.py   # undefined names in code paths the CPU suite cannot execute (GPU-only branches)
kungfu_b200/bin/kungfu-unit-tests
python -m pytest tests -x -q -m "not gpu"
scripts/sanitize.sh thread
echo "ci: all green"
