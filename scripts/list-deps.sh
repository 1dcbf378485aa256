#!/bin/sh
# What the build uses on this machine: compilers, CUDA toolkit, torch / NCCL versions, shared libraries of the built artefacts
# (reference: scripts/list-deps.sh).
cd "$(dirname "$0")/.."
echo "g++:   $(g++ --version | head -1)"
echo "nvcc:  $(nvcc --version 2>/dev/null | tail -2 | head -1)"
python - <<'PY'
import torch
print("torch:", torch.__version__, "cuda", torch.version.cuda, "nccl", ".".join(map(str, torch.cuda.nccl.version())) if torch.cuda.is_available() else "(no GPU)")
PY
for f in kungfu_b200/_kfrt*.so kungfu_b200/_kfcuda*.so kungfu_b200/bin/kungfu-run; do
    [ -e "$f" ] || continue
    echo "== $f"
    ldd "$f" | awk '{print "   " $1}' | sort -u
done
