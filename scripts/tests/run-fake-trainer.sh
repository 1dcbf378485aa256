#!/bin/sh
# Communication-only trainers, one back-end or all (reference: scripts/tests/run-fake-trainer.sh:98-137 <kungfu|mpi|nccl|go|tf|inproc|all>).
#   host    C ABI trainer under kungfu-run (per tensor / fused / async), np 1..4
#   inproc  workers as threads of one process: simulated shared-memory all-reduce (upper bound) and real peers over loopback
#   gpu     sm_100a collective kernels next to NCCL, per tensor and fused, incl. the 1-90 MiB simple_test (needs >= 2 GPUs)
set -e
. "$(dirname "$0")/common.sh"
cd "$ROOT"
BIN="$ROOT/kungfu_b200/bin"
what=${1:-all}
model=${2:-resnet50}
run_host() {
    for np in 1 2 3 4; do
        for mode in bench bench-fused bench-async; do prun $np "$BIN/kungfu-fake-agent" $mode "$model"; done
    done
}
run_inproc() {
    for np in 1 2 4; do srun "$BIN/kungfu-fake-inproc-trainer" -np $np -model "$model" -steps 5; done
    srun "$BIN/kungfu-fake-inproc-trainer" -np 4 -model "$model" -fuse -steps 5
}
run_gpu() {
    n=$(python -c 'import torch; print(torch.cuda.device_count())')
    if [ "$n" -lt 2 ]; then echo "gpu fake trainer skipped: $n GPU(s)"; return 0; fi
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29812 \
        benchmarks/fake_trainer.py --model "$model"
}
case $what in
host) run_host ;;
inproc) run_inproc ;;
gpu | nccl) run_gpu ;;
all) run_host; run_inproc; run_gpu ;;
*) echo "usage: $0 <host|inproc|gpu|all> [resnet50|vgg16|bert|slp]"; exit 2 ;;
esac
