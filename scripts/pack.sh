#!/bin/sh
# Source tarball of the tree without build products (reference: scripts/pack.sh).
set -e
cd "$(dirname "$0")/.."
name=$(basename "$(pwd)")
out=${1:-../$name.tar.gz}
tar --exclude='.git' --exclude='gpurun_out' --exclude='baseline/_ref' --exclude='*.so' --exclude='*.o' --exclude='__pycache__' \
    --exclude='kungfu_b200/build*' --exclude='kungfu_b200/bin*' --exclude='kungfu_b200/lib*' -czf "$out" -C .. "$name"
du -h "$out"
