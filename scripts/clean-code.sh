#!/bin/sh
# Format / sanity pass (reference: scripts/clean-code.sh): clang-format for C++/CUDA when available, byte-compile all Python,
# strip trailing whitespace. Build products are left alone.
set -e
cd "$(dirname "$0")/.."
if command -v clang-format >/dev/null 2>&1; then
    find kungfu_b200/csrc \( -name '*.cpp' -o -name '*.hpp' -o -name '*.h' -o -name '*.cu' -o -name '*.cuh' \) -print0 | xargs -0 clang-format -i
else
    echo "clang-format not installed: skipping C++/CUDA formatting"
fi
python -m compileall -q kungfu_b200 kungfu tests benchmarks examples experimental bench.py __graft_entry__.py
git ls-files '*.py' '*.sh' '*.md' '*.cu' '*.cuh' '*.cpp' '*.hpp' '*.h' | xargs sed -i 's/[ \t]*$//'
echo ok
