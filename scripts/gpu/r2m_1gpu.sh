#!/bin/bash
# smoke() with the transformer / GEMM kernels added (the last GPU seconds of round 2)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 80 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2m_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/r2m_smoke.log | cut -c1-400
