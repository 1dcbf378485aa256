#!/bin/sh
# Host-plane all-reduce benchmark over every strategy (reference: scripts/tests/run-allreduce-benchmark.sh), then the GPU sweep
# when GPUs are visible.
set -e
. "$(dirname "$0")/common.sh"
cd "$ROOT"
for strategy in STAR RING CLIQUE TREE BINARY_TREE BINARY_TREE_STAR MULTI_BINARY_TREE_STAR AUTO; do
    prun 4 -strategy $strategy "$ROOT/kungfu_b200/bin/kungfu-bench-allreduce" -model resnet50 -mode par -epochs 3 -warmup 1
    prun 4 -strategy $strategy "$ROOT/kungfu_b200/bin/kungfu-bench-allreduce" -model resnet50 -fuse -epochs 3 -warmup 1
    prun 4 -strategy $strategy python -m kungfu_b200.benchmarks --method CPU --warmup-steps 1 --steps 3
done
if python -c "import torch,sys; sys.exit(0 if torch.cuda.device_count() >= 2 else 1)"; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$(python -c 'import torch; print(torch.cuda.device_count())')" \
        --master-addr 127.0.0.1 --master-port 29811 benchmarks/allreduce_bw.py --nccl
fi
