#!/bin/sh
# Fetch CIFAR-10 / CIFAR-100 (python pickles) into $DATA_DIR/cifar, where `kungfu_b200.helpers.cifar` looks (reference:
# scripts/download-cifar.sh). Needs network access.
set -e
dir=${DATA_DIR:-$HOME/var/data}/cifar
mkdir -p "$dir" && cd "$dir"
prefix=${CIFAR_MIRROR:-https://www.cs.toronto.edu/~kriz}
for f in cifar-10-python.tar.gz cifar-100-python.tar.gz; do
    [ -f "$f" ] || { if command -v curl >/dev/null 2>&1; then curl -fsSLO "$prefix/$f"; else wget -q "$prefix/$f"; fi; }
    tar -xzf "$f"
done
ls -l "$dir"
