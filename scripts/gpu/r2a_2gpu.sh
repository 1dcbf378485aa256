#!/bin/bash
# Round-2 first GPU pass (2 GPUs): smoke, N=2 correctness of the split-phase kernels, A/B of gated vs single-kernel vs NCCL arm.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export KUNGFU_DEVICE_TIMEOUT_MS=8000
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpus.csv 2>&1
( timeout 400 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2a_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/r2a_summary.txt
( timeout 700 $TR --nproc-per-node 2 --master-port 29611 tests/mp_gpu_worker.py allreduce bcast ssgd extra ) > gpurun_out/r2a_mp2.log 2>&1
echo "mp2 rc=$? $(grep -o '"fails": [0-9]*' gpurun_out/r2a_mp2.log | tail -1)" | tee -a gpurun_out/r2a_summary.txt
run() { # name gpus extra-env... -- bench args
  local name=$1 n=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  if [ "$n" = 1 ]; then
    ( env "${envs[@]}" timeout 400 python bench.py --gpus 1 "$@" ) > gpurun_out/r2a_$name.log 2>&1
  else
    ( env "${envs[@]}" timeout 400 $TR --nproc-per-node $n --master-port $((29620 + RANDOM % 300)) bench.py --gpus $n "$@" ) > gpurun_out/r2a_$name.log 2>&1
  fi
  echo "$name rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2a_$name.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2a_$name.log | head -1)" | tee -a gpurun_out/r2a_summary.txt
}
run ours1 1 X=1 -- --steps 30 --warmup 5 --no-e2e
run ours2_gated 2 KUNGFU_GATED=1 -- --steps 30 --warmup 5 --no-e2e --trace gpurun_out/trace_ours2_gated
run ours2_single 2 KUNGFU_GATED=0 -- --steps 30 --warmup 5 --no-e2e --trace gpurun_out/trace_ours2_single
run ours2_gated_b8 2 KUNGFU_GATED=1 KUNGFU_COMM_BLOCKS=8 -- --steps 30 --warmup 5 --no-e2e
run nccl2 2 X=1 -- --impl nccl --steps 30 --warmup 5 --no-e2e --trace gpurun_out/trace_nccl2
run sma2_ours 2 X=1 -- --optimizer sma --steps 30 --warmup 5 --no-e2e
run sma2_nccl 2 X=1 -- --optimizer sma --impl nccl --steps 30 --warmup 5 --no-e2e
cat gpurun_out/r2a_summary.txt
