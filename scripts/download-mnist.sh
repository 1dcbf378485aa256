#!/bin/sh
# Fetch the four MNIST idx files into $DATA_DIR/mnist (default ~/var/data/mnist), where `kungfu_b200.helpers.mnist.load_datasets` looks
# (reference: scripts/download-mnist.sh). Needs network access; the examples fall back to MNIST-shaped synthetic data without it.
set -e
cd "$(dirname "$0")"
. ./utils/ensure_file.sh
dir=${DATA_DIR:-$HOME/var/data}/mnist
mkdir -p "$dir" && cd "$dir"
prefix=${MNIST_MIRROR:-https://storage.googleapis.com/cvdf-datasets/mnist}
ensure_file 6c95f4b05d2bf285e1bfb0e7960c31bd3b3f8a7d train-images-idx3-ubyte.gz "$prefix"
ensure_file 2a80914081dc54586dbdf242f9805a6b8d2a15fc train-labels-idx1-ubyte.gz "$prefix"
ensure_file c3a25af1f52dad7f726cce8cacb138654b760d48 t10k-images-idx3-ubyte.gz "$prefix"
ensure_file 763e7fa3757d93b0cdec073cef058b2004252c17 t10k-labels-idx1-ubyte.gz "$prefix"
for f in *.gz; do [ -f "${f%.gz}" ] || gzip -dk "$f"; done
ls -l "$dir"
