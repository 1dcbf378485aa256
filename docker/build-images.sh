#!/bin/sh
# Build the images of this directory from the repository root (reference: docker/build-images.sh).
#   docker/build-images.sh [cpu|cuda|example|all] [tag-prefix]
set -e
cd "$(dirname "$0")/.."
what=${1:-all}
prefix=${2:-kungfu-b200}
build() { docker build --rm -t "$prefix:$1" -f "docker/Dockerfile.$1" .; }
case $what in
cpu | cuda) build "$what" ;;
example) build cpu && docker build --rm -t "$prefix:example" --build-arg BASE="$prefix:cpu" -f docker/Dockerfile.example . ;;
all) build cpu && build cuda && docker build --rm -t "$prefix:example" --build-arg BASE="$prefix:cpu" -f docker/Dockerfile.example . ;;
*) echo "usage: $0 [cpu|cuda|example|all] [tag-prefix]" >&2; exit 2 ;;
esac
